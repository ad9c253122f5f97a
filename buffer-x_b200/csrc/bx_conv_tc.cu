// bx_conv_tc.cu -- a8/a11 convolution stacks on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same implicit GEMM as bx_conv.cu (rows = (sample, output position), cols = Cout, K = taps*Cin, padding
// geometry folded into the loader), but the inner product runs as tcgen05.mma kind::tf32 with fp32
// accumulators in tensor memory.  fp32-grade accuracy (descriptor parity 1e-4 rel) comes from the
// 3xTF32 split: x = hi + lo with hi = cvt.rna.tf32(x), lo = x - hi (exact);  a*b ~= ah*bh + ah*bl + al*bh
// (the dropped al*bl term is ~2^-22 relative).  Weights are split on the host, activations in the loader.
//
// CTA = 256 threads = 256 GEMM rows = two M=128 MMA tiles sharing one B tile; N = NT (32/64/128) columns
// -> 2*NT TMEM columns.  A stage is 16 input channels of one tap: two K=8 steps, each with hi and lo
// operand images written by the loader threads straight into the canonical K-major no-swizzle UMMA
// layout (core matrix = 8 rows x 16 bytes; thread r writes row r with 16-byte STS -> conflict-free):
//     A image  [kunit(2)][row(256)][16 B]           LBO = 4096 B, SBO = 128 B
//     B image  [kunit(2)][n(NT)][16 B]              LBO = NT*16 B, SBO = 128 B   (pre-arranged on the host)
// Two stages are double-buffered; tcgen05.commit on an mbarrier per stage tells the loaders when the tensor
// core has finished reading a stage.  The k order is (16-channel chunk outer, tap inner) so that the 9/27
// taps of a chunk re-read the same activations from L1.  Epilogue: tcgen05.ld 32x32b -> bias (+ReLU) ->
// coalesced stores of out[n][co][pos].
#include "bx_common.cuh"

namespace {

struct ConvTcParams {
    const float *in, *w, *bias;
    float *out;
    int n;
    const int *d_n;
    int Cin, Cout, D, H, W, kd, kh, kw, relu;
    int S_in, S_out, OD, OH, OW, T;
    const float *equi_s, *equi_t;
    const int *s_mids, *t_mids;
};

constexpr int TC_THREADS = 256;
constexpr int TC_BM = 256;
constexpr int A_STAGE_BYTES = 2 * 2 * 2 * TC_BM * 16;  // [kstep][split][kunit][row][16B] = 32 KB

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mma_commit(uint32_t bar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

#define TMEM_LD32(taddr, v)                                                                                    \
    asm volatile(                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, " \
        "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"    \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),      \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),             \
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),           \
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),           \
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                                \
        : "r"(taddr)                                                                                           \
        : "memory")

template <int GEOM, int NT>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(const ConvTcParams p) {
    constexpr int B_STAGE_BYTES = 2 * 2 * 2 * NT * 16;  // [kstep][split][kunit][n][16B]
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int TMEM_COLS = (2 * NT < 32) ? 32 : 2 * NT;  // 64 / 128 / 256: powers of two
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmem_base_s;

    const int n_samples = p.d_n ? *p.d_n : p.n;
    const long long Mtotal = (long long)n_samples * p.S_out;
    const long long row0 = (long long)blockIdx.x * TC_BM;
    if (row0 >= Mtotal) return;  // uniform per CTA, before any barrier / TMEM allocation
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    // ---- loader geometry: this thread owns GEMM row (row0 + tid) ---------------------------------------
    const long long lm = row0 + tid;
    const bool lvalid = lm < Mtotal;
    int ln = 0, oz = 0, oy = 0, ox = 0;
    if (lvalid) {
        ln = (int)(lm / p.S_out);
        const int pos = (int)(lm - (long long)ln * p.S_out);
        if (GEOM == BX_GEOM_CYL3D || GEOM == BX_GEOM_CYL2D) {
            oy = pos / 20;
            ox = pos - oy * 20;
        } else {
            oz = pos / (p.OH * p.OW);
            const int rem = pos - oz * (p.OH * p.OW);
            oy = rem / p.OW;
            ox = rem - oy * p.OW;
        }
    }
    const float *pa = p.in, *pb = nullptr;
    if (GEOM == BX_GEOM_COSTVOL) {
        if (lvalid) {
            pa = p.equi_s + (size_t)p.s_mids[ln] * 32 * 140;
            pb = p.equi_t + (size_t)p.t_mids[ln] * 32 * 140;
        }
    } else {
        pa = p.in + (size_t)ln * p.Cin * p.S_in;
    }
    const int cstride = (GEOM == BX_GEOM_COSTVOL) ? 140 : p.S_in;
    const int n_iters = (p.Cin / 16) * p.T;  // stage = (16-channel chunk, tap); chunk outer, tap inner

    float a_reg[16];
    float4 b_reg[(B_STAGE_BYTES / 16 + TC_THREADS - 1) / TC_THREADS];
    constexpr int B_VEC = B_STAGE_BYTES / 16;                       // 16-byte vectors per stage
    constexpr int B_PER_T = (B_VEC + TC_THREADS - 1) / TC_THREADS;  // 4 / 2 / 1

    auto load_stage = [&](int it) {
        const int chunk = it / p.T, t = it - chunk * p.T;
        const int dz = t / (p.kh * p.kw);
        const int r2 = t - dz * (p.kh * p.kw);
        const int dy = r2 / p.kw;
        const int dx = r2 - dy * p.kw;
        int offA = 0, offB = 0;
        bool ok = lvalid;
        if (GEOM == BX_GEOM_CYL3D || GEOM == BX_GEOM_CYL2D) {
            const int yy = oy + dy - 1;
            int xx = ox + dx - 1;
            xx = xx < 0 ? xx + 20 : (xx >= 20 ? xx - 20 : xx);
            ok = lvalid && yy >= 0 && yy < 7;
            offA = dz * 140 + yy * 20 + xx;
        } else if (GEOM == BX_GEOM_VALID3D) {
            offA = ((oz + dz) * p.H + (oy + dy)) * p.W + (ox + dx);
        } else {
            const int nn = oz + dz, kk = oy + dy, ll = ox + dx;
            int sh = ll - nn;
            sh = sh < 0 ? sh + 20 : sh;
            offA = (1 + kk) * 20 + sh;
            offB = (1 + kk) * 20 + ll;
        }
        const int ci0 = chunk * 16;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float v = 0.0f;
            if (ok) {
                const size_t o = (size_t)(ci0 + kk) * cstride;
                if (GEOM == BX_GEOM_COSTVOL) v = pa[o + offA] - pb[o + offB];
                else v = __ldg(pa + o + offA);
            }
            a_reg[kk] = v;
        }
        const float4 *wsrc = reinterpret_cast<const float4 *>(p.w) + (size_t)it * B_VEC;
#pragma unroll
        for (int j = 0; j < B_PER_T; ++j) {
            const int v = tid + j * TC_THREADS;
            if (v < B_VEC) b_reg[j] = __ldg(wsrc + v);
        }
    };
    auto store_stage = [&](int s) {
        unsigned char *As = smem + (size_t)s * STAGE_BYTES;
        unsigned char *Bs = As + A_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int ku = 0; ku < 2; ++ku) {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x = a_reg[ks * 8 + ku * 4 + j];
                    hi[j] = tf32_rna(x);
                    lo[j] = x - hi[j];
                }
                // [kstep][split][kunit][row][16B]
                *reinterpret_cast<float4 *>(As + ((size_t)((ks * 2 + 0) * 2 + ku) * TC_BM + tid) * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(As + ((size_t)((ks * 2 + 1) * 2 + ku) * TC_BM + tid) * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
#pragma unroll
        for (int j = 0; j < B_PER_T; ++j) {
            const int v = tid + j * TC_THREADS;
            if (v < B_VEC) *reinterpret_cast<float4 *>(Bs + (size_t)v * 16) = b_reg[j];
        }
    };

    // instruction descriptor: D=F32, A=B=TF32, both K-major, N = NT, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);

    load_stage(0);
    for (int it = 0; it < n_iters; ++it) {
        const int s = it & 1;
        if (it >= 2) mbar_wait(s ? bar1 : bar0, (uint32_t)(((it >> 1) - 1) & 1));  // MMAs of iteration it-2 have read stage s
        store_stage(s);
        if (it + 1 < n_iters) load_stage(it + 1);                                 // global loads in flight across the barrier
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");             // generic-proxy stores -> async proxy (tensor core)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_base = smem_base + (uint32_t)s * STAGE_BYTES;
            const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint64_t bh = make_desc(b_base + (uint32_t)((ks * 2 + 0) * 2) * NT * 16, NT * 16, 128);
                const uint64_t bl = make_desc(b_base + (uint32_t)((ks * 2 + 1) * 2) * NT * 16, NT * 16, 128);
#pragma unroll
                for (int tile = 0; tile < 2; ++tile) {
                    const uint64_t ah = make_desc(a_base + (uint32_t)((ks * 2 + 0) * 2) * TC_BM * 16 + tile * 2048, TC_BM * 16, 128);
                    const uint64_t al = make_desc(a_base + (uint32_t)((ks * 2 + 1) * 2) * TC_BM * 16 + tile * 2048, TC_BM * 16, 128);
                    const uint32_t d = tmem_base + (uint32_t)tile * NT;
                    const uint32_t first = (it == 0 && ks == 0) ? 0u : 1u;
                    mma_tf32(d, al, bh, IDESC, first);   // small terms first, the dominant hi*hi product last
                    mma_tf32(d, ah, bl, IDESC, 1u);
                    mma_tf32(d, ah, bh, IDESC, 1u);
                }
            }
            mma_commit(s ? bar1 : bar0);
        }
    }
    // ---- wait for the last MMAs (the commit of the final iteration covers everything issued before it) ----
    {
        const int last = n_iters - 1;
        mbar_wait((last & 1) ? bar1 : bar0, (uint32_t)((last >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- epilogue: thread = row (warp w reads TMEM lanes 32*(w%4).., accumulator tile w/4) ---------------
    {
        const int tile = warp >> 2, q = warp & 3;
        const long long m = row0 + tid;  // tid == tile*128 + q*32 + lane
        int n = 0, pos = 0;
        if (m < Mtotal) {
            n = (int)(m / p.S_out);
            pos = (int)(m - (long long)n * p.S_out);
        }
        float *o = p.out + (size_t)n * p.Cout * p.S_out + pos;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * NT + c0);
            TMEM_LD32(taddr, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < Mtotal) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = c0 + j;
                    if (co < p.Cout) {
                        float r = __uint_as_float(v[j]) + __ldg(p.bias + co);
                        if (p.relu) r = fmaxf(r, 0.0f);
                        o[(size_t)co * p.S_out] = r;
                    }
                }
            }
        }
        (void)lane;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int GEOM, int NT>
int launch_tc(const ConvTcParams &p, int max_n, cudaStream_t st) {
    const long long maxM = (long long)max_n * p.S_out;
    const unsigned gx = (unsigned)((maxM + TC_BM - 1) / TC_BM);
    if (gx == 0) return BX_OK;
    constexpr int smem = 2 * (A_STAGE_BYTES + 2 * 2 * 2 * NT * 16);
    static bool attr_done = false;
    if (!attr_done) {
        BX_CUDA(cudaFuncSetAttribute(conv_tc_kernel<GEOM, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
    }
    conv_tc_kernel<GEOM, NT><<<gx, TC_THREADS, smem, st>>>(p);
    BX_LAUNCH_CHECK();
    return BX_OK;
}

template <int GEOM>
int dispatch_nt(const ConvTcParams &p, int max_n, cudaStream_t st) {
    if (p.Cout > 64) return launch_tc<GEOM, 128>(p, max_n, st);
    if (p.Cout > 32) return launch_tc<GEOM, 64>(p, max_n, st);
    return launch_tc<GEOM, 32>(p, max_n, st);
}

}  // namespace

BX_API int bx_conv_tc_ntile(int Cout) { return Cout > 64 ? 128 : (Cout > 32 ? 64 : 32); }

BX_API int bx_conv_layer_tc(int geom, const float *in, const float *w_tc, const float *bias, float *out, int n,
                            const int32_t *d_n, int Cin, int Cout, int D, int H, int W, int kd, int kh, int kw, int relu,
                            const float *equi_s, const float *equi_t, const int32_t *s_mids, const int32_t *t_mids,
                            void *stream) {
    BX_REQUIRE(w_tc && bias && out, "bx_conv_layer_tc: null pointer");
    BX_REQUIRE(n >= 0 && Cin >= 16 && Cin % 16 == 0 && Cout >= 1 && Cout <= 128, "bx_conv_layer_tc: bad channels Cin=%d Cout=%d", Cin, Cout);
    BX_REQUIRE((reinterpret_cast<uintptr_t>(w_tc) & 15) == 0, "bx_conv_layer_tc: weights must be 16-byte aligned");
    ConvTcParams p = {};
    p.in = in; p.w = w_tc; p.bias = bias; p.out = out; p.n = n; p.d_n = d_n;
    p.Cin = Cin; p.Cout = Cout; p.D = D; p.H = H; p.W = W; p.kd = kd; p.kh = kh; p.kw = kw; p.relu = relu;
    p.equi_s = equi_s; p.equi_t = equi_t; p.s_mids = s_mids; p.t_mids = t_mids;
    p.T = kd * kh * kw;
    cudaStream_t st = bx_stream(stream);
    switch (geom) {
        case BX_GEOM_CYL3D:
            BX_REQUIRE(in && D == 3 && H == 7 && W == 20 && kd == 3 && kh == 3 && kw == 3, "bx_conv_layer_tc: CYL3D expects [C,3,7,20], k=3x3x3");
            p.S_in = 420; p.S_out = 140; p.OD = 1; p.OH = 7; p.OW = 20;
            return dispatch_nt<BX_GEOM_CYL3D>(p, n, st);
        case BX_GEOM_CYL2D:
            BX_REQUIRE(in && D == 1 && H == 7 && W == 20 && kd == 1 && kh == 3 && kw == 3, "bx_conv_layer_tc: CYL2D expects [C,7,20], k=3x3");
            p.S_in = 140; p.S_out = 140; p.OD = 1; p.OH = 7; p.OW = 20;
            return dispatch_nt<BX_GEOM_CYL2D>(p, n, st);
        case BX_GEOM_VALID3D:
            BX_REQUIRE(in && D >= kd && H >= kh && W >= kw && kd >= 1 && kh >= 1 && kw >= 1, "bx_conv_layer_tc: VALID3D kernel larger than input");
            p.OD = D - kd + 1; p.OH = H - kh + 1; p.OW = W - kw + 1;
            p.S_in = D * H * W; p.S_out = p.OD * p.OH * p.OW;
            return dispatch_nt<BX_GEOM_VALID3D>(p, n, st);
        case BX_GEOM_COSTVOL:
            BX_REQUIRE(equi_s && equi_t && s_mids && t_mids, "bx_conv_layer_tc: COSTVOL needs equi maps and match lists");
            BX_REQUIRE(Cin == 32 && D == 20 && H == 5 && W == 20 && kd == 3 && kh == 3 && kw == 3, "bx_conv_layer_tc: COSTVOL expects the [32,20,5,20] volume, k=3x3x3");
            p.OD = 18; p.OH = 3; p.OW = 18; p.S_in = 2000; p.S_out = 972;
            return dispatch_nt<BX_GEOM_COSTVOL>(p, n, st);
        default:
            bx_set_error("bx_conv_layer_tc: unknown geometry %d", geom);
            return BX_ERR_INVALID_ARG;
    }
}
