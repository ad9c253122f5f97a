// bx_conv_tc.cu -- a8/a11 convolution stacks on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same implicit GEMM as bx_conv.cu (rows = (sample, output position), cols = Cout, K = taps*Cin, padding
// geometry folded into the loader), but the inner product runs as tcgen05.mma kind::tf32 with fp32
// accumulators in tensor memory.  fp32-grade accuracy (descriptor parity 1e-4 rel) comes from the
// 3xTF32 split  x = hi + lo  (hi = x with the 13 low mantissa bits cleared, lo = x - hi, exact):
//     a*b ~= ah*bh + ah*bl + al*bh          (dropped al*bl ~ 2^-20 relative)
// The tensor core accumulates with truncation, one truncation per tcgen05.mma; the small cross terms
// therefore get their OWN accumulator (their truncation error is 2^-11 smaller) and only the ah*bh
// products go through the main one -- 3x fewer truncations on the value that matters.
//
// Warp-specialised CTA (544 threads, 1 CTA / SM, 256 GEMM rows = two M=128 tiles sharing each B tile):
//   warps 0-15  loaders : thread -> (row = t & 255, k-step = t >> 8).  Per stage (16 input channels of one
//               tap) a thread fetches its 8 activations (L1-resident across the taps of a chunk: the k
//               order is chunk-outer / tap-inner), splits them and writes hi/lo with 16-byte STS straight
//               into the canonical K-major no-swizzle UMMA layout (core matrix = 8 rows x 16 B):
//                   A image [kunit(2)][row(256)][16 B]   LBO = 4096 B, SBO = 128 B
//                   B image [kunit(2)][n(NT)][16 B]      LBO = NT*16 B, SBO = 128 B  (pre-arranged on the host)
//               then fence.proxy.async + mbarrier arrive on full[stage].
//               The B image of a stage is one contiguous 128*NT-byte block in global memory; loader thread 0
//               fetches it with ONE cp.async.bulk (UBLKCP) that signals the same full[stage] mbarrier
//               through its transaction count -- the weights never touch registers.
//   warp 16     MMA issuer: waits full[stage], issues 12 tcgen05.mma (2 k-steps x 2 tiles x 3 products),
//               tcgen05.commit -> empty[stage]; 4 stages in flight, no __syncthreads in the main loop.
//   epilogue    (loader warps) tcgen05.ld of main+cross accumulators -> bias (+ReLU) -> coalesced stores.
// Long reductions are cut into `nseg` segments: at a segment boundary the loader warps drain the
// accumulators into the (L2-resident) output tile with a properly rounded fp32 add and the next segment
// restarts from zero, which bounds the number of truncating accumulations per accumulator.
#include "bx_common.cuh"

namespace {

struct ConvTcParams {
    const float *in, *w, *bias;
    float *out;
    int n;
    const int *d_n;
    int Cin, Cout, D, H, W, kd, kh, kw, relu;
    int S_in, S_out, OD, OH, OW, T;
    int seg_len;  // stages per segment (>= n_iters: single segment)
    const float *equi_s, *equi_t;
    const int *s_mids, *t_mids;
};

constexpr int TC_LOADERS = 512;
constexpr int TC_THREADS = TC_LOADERS + 32;
constexpr int TC_BM = 256;
constexpr int TC_STAGES = 4;
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

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
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

#define TMEM_LD16(taddr, v)                                                                                   \
    asm volatile(                                                                                             \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, " \
        "%14, %15}, [%16];"                                                                                   \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
        : "r"(taddr)                                                                                          \
        : "memory")

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <int GEOM, int NT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvTcParams p) {
    constexpr int B_STAGE_BYTES = 2 * 2 * 2 * NT * 16;  // [kstep][split][kunit][n][16B]
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int TMEM_COLS = 4 * NT;  // two tiles x (main + cross-term accumulator): 128 / 256 / 512
    // barriers: full[4] (16 loader-warp arrivals + 1 expect_tx arrival), empty[4], segdone, accfree
    constexpr int BAR_EMPTY = TC_STAGES, BAR_SEGDONE = 2 * TC_STAGES, BAR_ACCFREE = 2 * TC_STAGES + 1;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * TC_STAGES + 2];
    __shared__ uint32_t tmem_base_s;

    const int n_samples = p.d_n ? *p.d_n : p.n;
    const long long Mtotal = (long long)n_samples * p.S_out;
    const long long row0 = (long long)blockIdx.x * TC_BM;
    if (row0 >= Mtotal) return;  // uniform per CTA, before any barrier / TMEM allocation
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n_iters = (p.Cin / 16) * p.T;  // stage = (16-channel chunk, tap); chunk outer, tap inner
    const int seg_len = p.seg_len;

    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), TC_LOADERS / 32 + 1);
            mbar_init(smem_u32(&bars[BAR_EMPTY + s]), 1);
        }
        mbar_init(smem_u32(&bars[BAR_SEGDONE]), 1);
        mbar_init(smem_u32(&bars[BAR_ACCFREE]), TC_LOADERS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(&bars[0]);

    if (warp < 16) {
        // =========================== loaders ===========================================================
        const int row = tid & 255, ks = tid >> 8;
        const long long lm = row0 + row;
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
        const int cstride = (GEOM == BX_GEOM_CYL2D || GEOM == BX_GEOM_COSTVOL) ? 140 : (GEOM == BX_GEOM_CYL3D ? 420 : p.S_in);
        // incremental (chunk, tap) counters: no integer division in the loop
        int chunk = 0, t = 0, dz = 0, dy = 0, dx = 0;
        float a_reg[8];

        auto load_stage = [&]() {
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
            } else {  // COSTVOL: value(c, n, k, l) = d1[c][1+k][(l-n) mod 20] - d2[c][1+k][l]
                const int nn = oz + dz, kk = oy + dy, ll = ox + dx;
                int sh = ll - nn;
                sh = sh < 0 ? sh + 20 : sh;
                offA = (1 + kk) * 20 + sh;
                offB = (1 + kk) * 20 + ll;
            }
            const float *src = pa + (size_t)(chunk * 16 + ks * 8) * cstride + offA;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                float v = 0.0f;
                if (ok) {
                    if (GEOM == BX_GEOM_COSTVOL) v = src[kk * cstride] - pb[(size_t)(chunk * 16 + ks * 8 + kk) * 140 + offB];
                    else v = __ldg(src + kk * cstride);
                }
                a_reg[kk] = v;
            }
            // advance to the next (chunk, tap)
            ++t;
            if (++dx == p.kw) {
                dx = 0;
                if (++dy == p.kh) { dy = 0; ++dz; }
            }
            if (t == p.T) { t = 0; dz = 0; dy = 0; dx = 0; ++chunk; }
        };
        auto store_stage = [&](int s) {
            unsigned char *As = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
            for (int ku = 0; ku < 2; ++ku) {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x = a_reg[ku * 4 + j];
                    hi[j] = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
                    lo[j] = x - hi[j];
                }
                // [kstep][split][kunit][row][16B]
                *reinterpret_cast<float4 *>(As + ((size_t)((ks * 2 + 0) * 2 + ku) * TC_BM + row) * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(As + ((size_t)((ks * 2 + 1) * 2 + ku) * TC_BM + row) * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        };
        // accumulators -> output tile.  mode 0: out = acc (first drain); 1: out += acc; 2: final (bias, ReLU)
        const int eq = warp & 3, etile = (warp >> 2) & 1, ehalf = warp >> 3;
        const long long em = row0 + etile * 128 + eq * 32 + (tid & 31);
        int en = 0, epos = 0;
        if (em < Mtotal) {
            en = (int)(em / p.S_out);
            epos = (int)(em - (long long)en * p.S_out);
        }
        float *eo = p.out + (size_t)en * p.Cout * p.S_out + epos;
        auto drain = [&](int mode, bool have_prev) {
            const uint32_t lane_base = (uint32_t)(eq * 32) << 16;
#pragma unroll 1
            for (int c0 = ehalf * 16; c0 < NT; c0 += 32) {
                uint32_t v[16], u[16];
                TMEM_LD16(tmem_base + lane_base + (uint32_t)(etile * NT + c0), v);           // main accumulator
                TMEM_LD16(tmem_base + lane_base + (uint32_t)(2 * NT + etile * NT + c0), u);  // cross terms
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (em < Mtotal) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int co = c0 + j;
                        if (co < p.Cout) {
                            float r = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                            float *dst = eo + (size_t)co * p.S_out;
                            if (have_prev) r += *dst;
                            if (mode == 2) {
                                r += __ldg(p.bias + co);
                                if (p.relu) r = fmaxf(r, 0.0f);
                            }
                            *dst = r;
                        }
                    }
                }
            }
        };

        load_stage();
        int seg = 0;
        for (int it = 0; it < n_iters; ++it) {
            const int s = it & (TC_STAGES - 1);
            if (it > 0 && it - seg * seg_len == seg_len) {
                // segment boundary: every MMA of the finished segment has completed -> drain, then free the accumulators
                mbar_wait(bar_base + 8u * BAR_SEGDONE, (uint32_t)(seg & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                drain(1, seg > 0);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar_base + 8u * BAR_ACCFREE);
                ++seg;
            }
            const uint32_t use = (uint32_t)(it / TC_STAGES);  // how many times this stage slot has been filled before
            if (use > 0) mbar_wait(bar_base + 8u * (BAR_EMPTY + s), (use - 1) & 1);  // tensor core has drained the slot
            if (tid == 0) {  // weights of this stage: one bulk copy, completion counted on full[s]
                mbar_arrive_expect_tx(bar_base + 8u * s, (uint32_t)B_STAGE_BYTES);
                bulk_g2s(smem_base + (uint32_t)s * STAGE_BYTES + A_STAGE_BYTES,
                         reinterpret_cast<const unsigned char *>(p.w) + (size_t)it * B_STAGE_BYTES, (uint32_t)B_STAGE_BYTES, bar_base + 8u * s);
            }
            store_stage(s);
            if (it + 1 < n_iters) load_stage();                   // next stage's activations in flight
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async proxy
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_base + 8u * s);
        }
        // ---- final epilogue ------------------------------------------------------------------------------
        mbar_wait(bar_base + 8u * BAR_SEGDONE, (uint32_t)(seg & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        drain(2, seg > 0);
    } else {
        // =========================== MMA issuer (warp 16) ==============================================
        // instruction descriptor: D=F32, A=B=TF32, both K-major, N = NT, M = 128
        constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if ((tid & 31) == 0) {
            int seg = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int s = it & (TC_STAGES - 1);
                bool seg_first = (it == 0);
                if (it > 0 && it - seg * seg_len == seg_len) {
                    mbar_wait(bar_base + 8u * BAR_ACCFREE, (uint32_t)(seg & 1));  // loaders have drained the accumulators
                    ++seg;
                    seg_first = true;
                }
                mbar_wait(bar_base + 8u * s, (uint32_t)((it / TC_STAGES) & 1));
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
                        const uint32_t d_main = tmem_base + (uint32_t)(tile * NT);
                        const uint32_t d_cross = tmem_base + (uint32_t)(2 * NT + tile * NT);
                        const uint32_t acc = (seg_first && ks == 0) ? 0u : 1u;
                        mma_tf32(d_cross, al, bh, IDESC, acc);
                        mma_tf32(d_cross, ah, bl, IDESC, 1u);
                        mma_tf32(d_main, ah, bh, IDESC, acc);
                    }
                }
                mma_commit(bar_base + 8u * (BAR_EMPTY + s));   // slot s may be refilled once these MMAs have read it
                if (it == n_iters - 1 || (it + 1) - seg * seg_len == seg_len) mma_commit(bar_base + 8u * BAR_SEGDONE);
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 16) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int GEOM, int NT>
int launch_tc(const ConvTcParams &p, int max_n, cudaStream_t st) {
    const long long maxM = (long long)max_n * p.S_out;
    const unsigned gx = (unsigned)((maxM + TC_BM - 1) / TC_BM);
    if (gx == 0) return BX_OK;
    constexpr int smem = TC_STAGES * (A_STAGE_BYTES + 2 * 2 * 2 * NT * 16);
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
    {   // segments: at most ~24 stages (= 48 truncating accumulations) per accumulator
        const int n_iters = (Cin / 16) * p.T;
        const int nseg = (n_iters + 23) / 24;
        p.seg_len = (n_iters + nseg - 1) / nseg;
    }
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
