// bx_conv_tc.cu -- a8/a11 convolution stacks on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same implicit GEMM as bx_conv.cu (rows = (sample, output position), cols = Cout, K = taps*Cin, padding
// geometry folded into the loader), but the inner product runs as tcgen05.mma kind::tf32 with fp32
// accumulators in tensor memory.  fp32-grade accuracy (descriptor parity 1e-4 rel) needs two things:
//   1. the 3xTF32 split  x = hi + lo  (hi = x with the 13 low mantissa bits cleared -- the tensor core ignores them,
//      so the hi operand is x itself -- lo = x - hi, exact):
//          a*b ~= ah*bh + ah*bl + al*bh          (dropped al*bl ~ 2^-20 relative)
//   2. short accumulation chains: the tensor core accumulates with truncation, so the error of a chain grows
//      linearly with its length (measured on B200, K = 1152: 5x the fp32-FFMA error for one chain, below it
//      for chains of 12 MMAs; tools/tc_precision.py).  The ah*bh products therefore go to a PING-PONG pair
//      of TMEM accumulators that is cut every seg_len stages (default 6 = 12 MMAs); finished segments are added with
//      round-to-nearest into fp32 running sums held in the loader threads' registers while the tensor core
//      already fills the other accumulator.  The cross terms (2^-11 smaller) keep one long chain.
//
// Warp-specialised CTA (LG*128 + 64 threads, 128 GEMM rows = one M=128 tile, N = NT columns):
//   loader warps  LG groups of four warps; group g owns the stages it = g, g+LG, ... so a warp touches a barrier once
//               per LG stages.  thread -> row = t & 127.  Per owned stage (16 input channels of one tap; chunk-outer /
//               tap-inner order keeps a chunk's activations in L1 across its taps; tap geometry from a shared table)
//               a thread fetches its row's 16 activations with four 16-byte loads (channel-blocked activations
//               [n][C/4][position][4]), splits them and writes hi/lo with tcgen05.st (32x32b.x8) into the A ring in
//               TENSOR MEMORY (4 stages x 32 columns: [kstep][hi,lo][8 values]); the MMAs read A from TMEM (TS
//               form), which takes the activation operand off the shared-memory read port -- with A in shared
//               memory the kernel was bound by smem bandwidth (3 MMAs re-read the same 128x8 tile).
//               Every seg_len stages each loader warp drains its 32 lanes x NT/LG columns of the finished main
//               accumulator (tcgen05.ld) into its running sums and releases the accumulator (accfree barrier).
//   weight warp   streams the host-arranged B (weight) images
//                   [kstep][split][kunit][n(NT)][16 B]      K-major no-swizzle, LBO = NT*16 B, SBO = 128 B
//               through a shared-memory ring with cp.async.bulk + mbarrier transaction counts, 8-12 stages AHEAD of
//               the tensor core and independent of the A slots (the weights never touch registers and their L2
//               latency is off the slot turnaround path).
//   MMA warp      waits a_full[stage] (and b_full once per weight super-stage), issues 6 tcgen05.mma
//               (2 k-steps x {al*bh, ah*bl, ah*bh}), tcgen05.commit -> a_empty / b_empty; no __syncthreads in the
//               main loop.
//   epilogue    running sums (+ cross accumulator) + bias (+ReLU) -> 16-byte channel-blocked stores.
// Configurations (dispatch_nt): Cout 128 -> LG=4, ping-pong main + separate cross accumulator, 1 CTA/SM;
// Cout 64 -> LG=2, ping-pong accumulators that also take the cross terms (segments of 4 stages), 2 CTAs/SM;
// Cout <= 32 -> LG=2, ping-pong + separate cross, 2 CTAs/SM.  -DBX_TC_TRACE adds a clock64 stage timeline.
#include "bx_common.cuh"
#include "bx_tcgen05.cuh"

namespace {

struct ConvTcParams {
    const float *in, *w, *bias;
    float *out;
    int n;
    const int *d_n;
    int Cin, Cout, D, H, W, kd, kh, kw, relu;
    int S_in, S_out, OD, OH, OW, T;
    int seg_len;  // stages per main-accumulator segment
    long long *trace;   // -DBX_TC_TRACE builds only: clock64 stamps of one CTA
    const float *equi_s, *equi_t;
    const int *s_mids, *t_mids;
};

constexpr int TC_BM = 128;
constexpr int TC_STAGES = 4;      // A ring (tensor memory)
// B ring (shared memory): weights are prefetched independently of the A slots, TC_SB stages per bulk copy and per
// barrier ("super-stage"), TC_NBS super-stages.  Narrow layers: 4 x 3 (one mbarrier wait per four stages -- the MMA
// warp's per-stage latency is what bounds them); wide layers: 1 x 8.
template <int NT> struct BRing { static constexpr int SB = NT == 128 ? 1 : 4, NBS = NT == 128 ? 8 : 3; };
constexpr int TC_MAX_TAPS = 128;
constexpr int A_STAGE_COLS = 32;  // TMEM columns of one A stage: [kstep(2)][split(hi,lo)][8 tf32 values]

// LG = loader groups of 4 warps (group g fills the stages it = g, g+LG, ...); NSETS = main accumulators (2 = ping-pong,
// 1 = the tensor core waits for the drain).  Wide layers (NT = 128) run LG = 4, NSETS = 2, one CTA per SM (384 of the 512
// TMEM columns).  Narrow layers (NT <= 64) are dominated by per-tile latencies (pipeline fill, drains, epilogue), not
// by tensor time, so they run LG = 2 (288 threads) with at most 256 TMEM columns: TWO CTAs per SM overlap one tile's
// prologue / drain bubbles / epilogue with the other tile's MMAs.
template <int GEOM, int NT, int LG, int NSETS, int MINB, int XSEP>
__global__ void __launch_bounds__(LG * 128 + 64, MINB) conv_tc_kernel(const ConvTcParams p) {
    constexpr int TC_LOADERS = LG * 128;
    constexpr int MMA_WARP = LG * 4;
    constexpr int WGT_WARP = LG * 4 + 1;   // weight producer: streams the B images through the shared-memory ring
    constexpr int TC_SB = BRing<NT>::SB, TC_NBS = BRing<NT>::NBS, TC_BSTAGES = TC_SB * TC_NBS;
    constexpr int B_STAGE_BYTES = 2 * 2 * 2 * NT * 16;  // [kstep][split][kunit][n][16B]
    constexpr int STAGE_BYTES = B_STAGE_BYTES;      // shared memory holds only the weights; A lives in tensor memory
    constexpr int A_RING = (NSETS + XSEP) * NT;     // TMEM columns: main[0..NSETS), [cross], then the A ring
    constexpr int TMEM_NEED = A_RING + TC_STAGES * A_STAGE_COLS;
    constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
    static_assert(MINB == 1 || TMEM_NEED <= 256, "two CTAs per SM need <= 256 TMEM columns each");
    constexpr int CW = NT / LG;         // accumulator columns owned by one loader warp
    // barriers: full[ST] (4 loader-warp arrivals + 1 expect_tx arrival), empty[ST], segdone[2], accfree[2]
    constexpr int BAR_EMPTY = TC_STAGES, BAR_SEGDONE = 2 * TC_STAGES, BAR_ACCFREE = 2 * TC_STAGES + 2;
    constexpr int BAR_BFULL = 2 * TC_STAGES + 4, BAR_BEMPTY = BAR_BFULL + TC_NBS;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * TC_STAGES + 4 + 2 * TC_NBS];
    __shared__ uint32_t tmem_base_s;
    __shared__ int4 tap_tab[TC_MAX_TAPS];

    const int n_samples = p.d_n ? *p.d_n : p.n;
    const long long Mtotal = (long long)n_samples * p.S_out;
    const long long row0 = (long long)blockIdx.x * TC_BM;
    if (row0 >= Mtotal) return;  // uniform per CTA, before any barrier / TMEM allocation
    const int tid = threadIdx.x, warp = tid >> 5;
#ifdef BX_TC_TRACE
    if (p.trace && blockIdx.x == gridDim.x / 2 && tid == 0) p.trace[4000] = clock64();   // CTA start
#endif
    const int n_iters = (p.Cin / 16) * p.T;  // stage = (16-channel chunk, tap); chunk outer, tap inner
    const int G = XSEP ? p.seg_len : (p.seg_len < 4 ? p.seg_len : 4);   // merged cross terms: 6 MMAs per stage in one chain
    const int nseg = (n_iters + G - 1) / G;

    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < p.T) {   // tap geometry table: no counter arithmetic in the loader loop
        const int dz = tid / (p.kh * p.kw), r = tid - dz * (p.kh * p.kw), dy = r / p.kw, dx = r - dy * p.kw;
        int base = 0;
        if (GEOM == BX_GEOM_CYL3D || GEOM == BX_GEOM_CYL2D) base = dz * 140 + (dy - 1) * 20;
        else if (GEOM == BX_GEOM_VALID3D) base = (dz * p.H + dy) * p.W + dx;
        tap_tab[tid] = make_int4(dz, dy, dx, base);
    }
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 4);                // A full: the 4 warps of the owning group
            mbar_init(smem_u32(&bars[BAR_EMPTY + s]), 1);
        }
        for (int s = 0; s < TC_NBS; ++s) {
            mbar_init(smem_u32(&bars[BAR_BFULL + s]), 1);    // B full: the producer's expect_tx arrival + the bulk copy's bytes
            mbar_init(smem_u32(&bars[BAR_BEMPTY + s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bars[BAR_SEGDONE + s]), 1);
            mbar_init(smem_u32(&bars[BAR_ACCFREE + s]), TC_LOADERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base = tmem_base_s;
    uint32_t smem_base = smem_u32(smem);
    uint32_t bar_base = smem_u32(&bars[0]);
    // opaque to the compiler: keep the three bases in registers instead of re-deriving the shared-window address
    // (S2R SR_CgaCtaId + LEA chains) in front of every barrier operation of the loader loop
    asm volatile("" : "+r"(tmem_base), "+r"(smem_base), "+r"(bar_base));

    if (warp < MMA_WARP) {
        // =========================== loaders ===========================================================
        const int row = tid & 127, grp = tid >> 7;       // grp: which stages (it % LG == grp) this thread fills
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
        } else if (GEOM == BX_GEOM_COSTAB) {
            pa = p.equi_s + (size_t)ln * 32 * 60;     // A, channel-blocked [8][3*20][4]
            pb = p.equi_t + (size_t)ln * 32 * 54;     // B, channel-blocked [8][3*18][4]
        } else {
            pa = p.in + (size_t)ln * p.S_in * p.Cin;      // activations are channel-blocked: [n][Cin/4][position][4]
        }
        constexpr int cstride = 140;   // channel-first equivariant maps of the direct cost-volume loader
        // (chunk, tap) of the stage this group fills next; the tap geometry comes from the shared table
        int chunk = grp / p.T, t = grp - chunk * p.T;
        auto advance_lg = [&]() {
            t += LG;
            while (t >= p.T) { t -= p.T; ++chunk; }
        };
        const int oy20 = oy * 20;
        const int rowbase = (oz * p.H + oy) * p.W + ox;   // VALID3D
        float a_reg[16];

        auto load_stage = [&]() {                       // the stage the counters point at
            const int4 tp = tap_tab[t];                  // (dz, dy, dx, geometry-specific base offset)
            int offA = 0, offB = 0;
            bool ok = lvalid;
            if (GEOM == BX_GEOM_CYL3D || GEOM == BX_GEOM_CYL2D) {
                const int yy = oy + tp.y - 1;
                int xx = ox + tp.z - 1;
                xx = xx < 0 ? xx + 20 : (xx >= 20 ? xx - 20 : xx);
                ok = lvalid && (unsigned)yy < 7u;
                offA = tp.w + oy20 + xx;
            } else if (GEOM == BX_GEOM_VALID3D) {
                offA = rowbase + tp.w;
            } else if (GEOM == BX_GEOM_COSTVOL) {  // value(c, n, k, l) = d1[c][1+k][(l-n) mod 20] - d2[c][1+k][l]
                const int nn = oz + tp.x, kk = oy + tp.y, ll = ox + tp.z;
                int sh = ll - nn;
                sh = sh < 0 ? sh + 20 : sh;
                offA = (1 + kk) * 20 + sh;
                offB = (1 + kk) * 20 + ll;
            } else {  // COSTAB: value(c, n, k, l) = relu(A[c][k][(l-n) mod 20] - B[c][k][l])
                const int nn = oz + tp.x, kk = oy + tp.y, ll = ox + tp.z;
                int sh = ll - nn;
                sh = sh < 0 ? sh + 20 : sh;
                offA = kk * 20 + sh;
                offB = kk * 18 + ll;
            }
            const int c0 = chunk * 16;
            if (GEOM == BX_GEOM_COSTVOL) {
                const float *src = pa + (size_t)c0 * cstride + offA;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    float v = 0.0f;
                    if (ok) v = src[kk * cstride] - pb[(size_t)(c0 + kk) * 140 + offB];
                    a_reg[kk] = v;
                }
            } else if (GEOM == BX_GEOM_COSTAB) {   // relu(A - B) from the channel-blocked factors: 2 x four 16-byte loads
                const float4 *sa = reinterpret_cast<const float4 *>(pa) + (chunk * 4) * 60 + offA;
                const float4 *sb4 = reinterpret_cast<const float4 *>(pb) + (chunk * 4) * 54 + offB;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 va = make_float4(0.0f, 0.0f, 0.0f, 0.0f), vb = va;
                    if (ok) { va = __ldg(sa + q * 60); vb = __ldg(sb4 + q * 54); }
                    a_reg[4 * q] = fmaxf(va.x - vb.x, 0.0f); a_reg[4 * q + 1] = fmaxf(va.y - vb.y, 0.0f);
                    a_reg[4 * q + 2] = fmaxf(va.z - vb.z, 0.0f); a_reg[4 * q + 3] = fmaxf(va.w - vb.w, 0.0f);
                }
            } else {  // four 16-byte loads, one per group of 4 channels; a warp's 32 rows read 512 contiguous bytes each
                const float4 *src = reinterpret_cast<const float4 *>(pa) + (size_t)(chunk * 4) * p.S_in + offA;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (ok) v = __ldg(src + (size_t)q * p.S_in);
                    a_reg[4 * q] = v.x; a_reg[4 * q + 1] = v.y; a_reg[4 * q + 2] = v.z; a_reg[4 * q + 3] = v.w;
                }
            }
        };
        const uint32_t a_lane = (uint32_t)((warp & 3) * 32) << 16;   // this warp's TMEM lanes = its 32 GEMM rows
        // registers -> tensor memory: [kstep][hi,lo][8].  The hi operand is the fp32 value itself: kind::tf32 reads only the
        // upper 19 bits of a 32-bit operand (truncation -- verified by the 2e-5 layer tests, which fail if the hardware
        // rounded), so x and (x & 0xFFFFE000) are the same operand and lo = x - (x & 0xFFFFE000) stays exact.
        auto store_stage = [&](int s) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                float xr[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xr[j] = a_reg[ks * 8 + j];
                    lo[j] = xr[j] - __uint_as_float(__float_as_uint(xr[j]) & 0xFFFFE000u);
                }
                const uint32_t col = (uint32_t)(A_RING + s * A_STAGE_COLS + ks * 16);
                tmem_st8(tmem_base + a_lane + col, xr);
                tmem_st8(tmem_base + a_lane + col + 8, lo);
            }
        };

        // ---- accumulator ownership of this warp: TMEM lanes 32*(warp&3).., columns CW*(warp>>2).. ----------
        const int eq = warp & 3, ecs = warp >> 2;
        const uint32_t tm_lane = (uint32_t)(eq * 32) << 16;
        float run[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) run[j] = 0.0f;
        auto drain = [&](int j, bool release) {       // add finished segment j (main set j&1) into the running sums
            const int set = j % NSETS;
            mbar_wait(bar_base + 8u * (BAR_SEGDONE + set), (uint32_t)((j / NSETS) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[CW];
            tmem_ld<CW>(tmem_base + tm_lane + (uint32_t)(set * NT + ecs * CW), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < CW; ++c) run[c] += __uint_as_float(v[c]);
            if (release) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar_base + 8u * (BAR_ACCFREE + set));
            }
        };

        if (grp < n_iters) load_stage();
        int next_drain = 0;
#ifdef BX_TC_TRACE
        const bool tr = p.trace && blockIdx.x == gridDim.x / 2 && (tid & 127) == 0;
        long long *tb = p.trace + (size_t)grp * 3 * 64;
        int trk = 0;
#endif
        for (int it = grp; it < n_iters; it += LG) {
#ifdef BX_TC_TRACE
            if (tr && trk < 64) tb[trk * 3 + 0] = clock64();
#endif
            if (next_drain < nseg - 1 && it >= (next_drain + 1) * G + (G < TC_STAGES ? G : TC_STAGES)) {
                drain(next_drain, true);                  // its MMAs are several stages behind us: short wait
                ++next_drain;
            }
            const int s = it % TC_STAGES;
            const uint32_t use = (uint32_t)(it / TC_STAGES);   // how many times slot s has been filled before
            if (use > 0) mbar_wait(bar_base + 8u * (BAR_EMPTY + s), (use - 1) & 1);  // tensor core has read the slot
#ifdef BX_TC_TRACE
            if (tr && trk < 64) tb[trk * 3 + 1] = clock64();
#endif
            store_stage(s);                               // tcgen05.st issued (sources are read at issue) ...
            if (it + LG < n_iters) {                      // ... this group's next activations go in flight behind them ...
                advance_lg();
                load_stage();
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");      // ... and only then wait for the stores
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // tcgen05.st ordered before the arrive
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_base + 8u * s);
#ifdef BX_TC_TRACE
            if (tr && trk < 64) { tb[trk * 3 + 2] = clock64(); ++trk; }
#endif
        }
#ifdef BX_TC_TRACE
        if (tr) p.trace[4002] = clock64();               // loader group: main loop done
#endif
        while (next_drain < nseg) {                      // a set must still be released if a later segment reuses it
            drain(next_drain, next_drain + NSETS < nseg);
            ++next_drain;
        }
        // ---- epilogue: running sums + cross accumulator + bias (+ReLU) ------------------------------------
        {
            uint32_t u[CW];
            if (XSEP) {
                tmem_ld<CW>(tmem_base + tm_lane + (uint32_t)(NSETS * NT + ecs * CW), u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int c = 0; c < CW; ++c) u[c] = 0u;   // cross terms were accumulated with the main products
            }
            const long long em = row0 + eq * 32 + (tid & 31);
            if (em < Mtotal) {
                const int en = (int)(em / p.S_out);
                const int epos = (int)(em - (long long)en * p.S_out);
                // channel-blocked output [n][Cout/4][position][4]: one 16-byte store per group of 4 channels, coalesced
                // across the warp's 32 consecutive rows
                float4 *eo = reinterpret_cast<float4 *>(p.out) + ((size_t)en * (p.Cout >> 2) + ((ecs * CW) >> 2)) * p.S_out + epos;
#pragma unroll
                for (int c = 0; c < CW; c += 4) {
                    const int co = ecs * CW + c;
                    if (co < p.Cout) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + co));
                        float4 r;
                        r.x = (run[c] + __uint_as_float(u[c])) + b4.x;
                        r.y = (run[c + 1] + __uint_as_float(u[c + 1])) + b4.y;
                        r.z = (run[c + 2] + __uint_as_float(u[c + 2])) + b4.z;
                        r.w = (run[c + 3] + __uint_as_float(u[c + 3])) + b4.w;
                        if (p.relu) { r.x = fmaxf(r.x, 0.0f); r.y = fmaxf(r.y, 0.0f); r.z = fmaxf(r.z, 0.0f); r.w = fmaxf(r.w, 0.0f); }
                        eo[(size_t)(c >> 2) * p.S_out] = r;
                    }
                }
            }
        }
    } else if (warp == WGT_WARP) {
        // =========================== weight producer ====================================================
        // The weight image of stage `it` is one contiguous block; its order is known in advance, so the producer runs
        // up to TC_BSTAGES stages ahead of the tensor core, independent of the activation slots: the L2 -> shared
        // memory latency of the bulk copy (~1 us) is off the stage turnaround path.
        if ((tid & 31) == 0) {
            const int n_super = (n_iters + TC_SB - 1) / TC_SB;
            for (int q = 0; q < n_super; ++q) {
                const int sb = q % TC_NBS;
                const uint32_t useb = (uint32_t)(q / TC_NBS);
                if (useb > 0) mbar_wait(bar_base + 8u * (BAR_BEMPTY + sb), (useb - 1) & 1);
                const int nst = (n_iters - q * TC_SB) < TC_SB ? (n_iters - q * TC_SB) : TC_SB;
                const uint32_t bytes = (uint32_t)nst * (uint32_t)B_STAGE_BYTES;
                mbar_arrive_expect_tx(bar_base + 8u * (BAR_BFULL + sb), bytes);
                bulk_g2s(smem_base + (uint32_t)(sb * TC_SB) * STAGE_BYTES,
                         reinterpret_cast<const unsigned char *>(p.w) + (size_t)q * TC_SB * B_STAGE_BYTES, bytes, bar_base + 8u * (BAR_BFULL + sb));
            }
        }
        __syncwarp();
    } else {
        // =========================== MMA issuer ==========================================================
        // instruction descriptor: D=F32, A=B=TF32, both K-major, N = NT, M = 128
        constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1
        constexpr uint32_t B_LBO = ((uint32_t)(NT * 16) >> 4) << 16;
        constexpr uint32_t B_IMG = (2u * NT * 16) >> 4;                         // one (kstep, split) image of B, in 16-byte units
        const uint32_t leader = elect_leader();
        const uint32_t b0 = (smem_base >> 4) | B_LBO;                            // stage 0, kstep 0, hi
        const uint32_t d_cross = tmem_base + (uint32_t)(NSETS * NT);
        int s = 0, sb = 0, seg = 0, in_seg = 0;
        uint32_t use = 0, useb = 0;
#ifdef BX_TC_TRACE
        const bool trm = p.trace && blockIdx.x == gridDim.x / 2 && (tid & 31) == 0;
        long long *tm = p.trace + 4 * 3 * 64;
#endif
        for (int it = 0; it < n_iters; ++it) {
#ifdef BX_TC_TRACE
            if (trm && it < 128) tm[it * 3 + 0] = clock64();
#endif
            if (in_seg == 0 && seg >= NSETS) {
                // segment `seg` reuses main set seg % NSETS: segment seg-NSETS must have been drained
                mbar_wait(bar_base + 8u * (BAR_ACCFREE + (seg % NSETS)), (uint32_t)(((seg - NSETS) / NSETS) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (sb % TC_SB == 0)                                      // weights: one wait per TC_SB stages (usually long since there)
                mbar_wait(bar_base + 8u * (BAR_BFULL + sb / TC_SB), useb & 1);
#ifdef BX_TC_TRACE
            if (trm && it < 128) tm[it * 3 + 1] = clock64();
#endif
            mbar_wait(bar_base + 8u * s, use & 1);                    // activations
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef BX_TC_TRACE
            if (trm && it < 128) tm[it * 3 + 2] = clock64();
#endif
            const uint32_t so = (uint32_t)sb * (uint32_t)(STAGE_BYTES >> 4);
            const uint32_t d_main = tmem_base + (uint32_t)((seg % NSETS) * NT);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint32_t ah = tmem_base + (uint32_t)(A_RING + s * A_STAGE_COLS + ks * 16), al = ah + 8;
                const uint32_t bh = b0 + so + (uint32_t)(ks * 2 + 0) * B_IMG, bl = b0 + so + (uint32_t)(ks * 2 + 1) * B_IMG;
                if (XSEP) {
                    mma_tf32_ts(leader, d_cross, al, bh, DESC_HI, IDESC, (it == 0 && ks == 0) ? 0u : 1u);
                    mma_tf32_ts(leader, d_cross, ah, bl, DESC_HI, IDESC, 1u);
                    mma_tf32_ts(leader, d_main, ah, bh, DESC_HI, IDESC, (in_seg == 0 && ks == 0) ? 0u : 1u);
                } else {   // one accumulator per segment takes all three products (small terms first)
                    mma_tf32_ts(leader, d_main, al, bh, DESC_HI, IDESC, (in_seg == 0 && ks == 0) ? 0u : 1u);
                    mma_tf32_ts(leader, d_main, ah, bl, DESC_HI, IDESC, 1u);
                    mma_tf32_ts(leader, d_main, ah, bh, DESC_HI, IDESC, 1u);
                }
            }
            mma_commit(leader, bar_base + 8u * (BAR_EMPTY + s));   // slot s may be refilled once these MMAs have read it
            if (sb % TC_SB == TC_SB - 1 || it == n_iters - 1)
                mma_commit(leader, bar_base + 8u * (BAR_BEMPTY + sb / TC_SB));   // the whole super-stage has been read
            if (++s == TC_STAGES) { s = 0; ++use; }
            if (++sb == TC_BSTAGES) { sb = 0; ++useb; }
            if (++in_seg == G || it == n_iters - 1) {
                mma_commit(leader, bar_base + 8u * (BAR_SEGDONE + (seg % NSETS)));
                in_seg = 0;
                ++seg;
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
#ifdef BX_TC_TRACE
    if (p.trace && blockIdx.x == gridDim.x / 2 && tid == 0) p.trace[4001] = clock64();   // CTA end (before dealloc)
#endif
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int GEOM, int NT, int LG, int NSETS, int MINB, int XSEP>
int launch_tc(const ConvTcParams &p, int max_n, cudaStream_t st) {
    const long long maxM = (long long)max_n * p.S_out;
    const unsigned gx = (unsigned)((maxM + TC_BM - 1) / TC_BM);
    if (gx == 0) return BX_OK;
    constexpr int smem = BRing<NT>::SB * BRing<NT>::NBS * (2 * 2 * 2 * NT * 16);
    static BxPerDevice attr_done = {};
    if (bx_needs_attr(attr_done))
        BX_CUDA(cudaFuncSetAttribute(conv_tc_kernel<GEOM, NT, LG, NSETS, MINB, XSEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    conv_tc_kernel<GEOM, NT, LG, NSETS, MINB, XSEP><<<gx, LG * 128 + 64, smem, st>>>(p);
    BX_LAUNCH_CHECK();
    return BX_OK;
}

int g_tc_mode = -1;   // -1: read BX_TC_MODE once (debug / A-B switch): 0 default, 1 = every layer one CTA per SM (LG=4),
                      // 2 = narrow layers two CTAs per SM with 16 loader warps each (60 registers per thread)

template <int GEOM>
int dispatch_nt(const ConvTcParams &p, int max_n, cudaStream_t st) {
    if (g_tc_mode < 0) {
        const char *e = getenv("BX_TC_MODE");
        g_tc_mode = e ? atoi(e) : 0;
    }
    if (p.Cout > 64) return launch_tc<GEOM, 128, 4, 2, 1, 1>(p, max_n, st);
    if (g_tc_mode == 1) {           // every layer one CTA per SM
        if (p.Cout > 32) return launch_tc<GEOM, 64, 4, 2, 1, 1>(p, max_n, st);
        return launch_tc<GEOM, 32, 4, 2, 1, 1>(p, max_n, st);
    }
    if (g_tc_mode == 2) {           // narrow layers: two CTAs per SM with 16 loader warps each (56 registers per thread)
        if (p.Cout > 32) return launch_tc<GEOM, 64, 4, 2, 2, 0>(p, max_n, st);
        return launch_tc<GEOM, 32, 4, 2, 2, 1>(p, max_n, st);
    }
    if (g_tc_mode == 3) {           // Cout 64: single main accumulator + separate cross accumulator (drain bubble)
        if (p.Cout > 32) return launch_tc<GEOM, 64, 2, 1, 2, 1>(p, max_n, st);
        return launch_tc<GEOM, 32, 2, 2, 2, 1>(p, max_n, st);
    }
    if (p.Cout > 32) return launch_tc<GEOM, 64, 2, 2, 2, 0>(p, max_n, st);
    return launch_tc<GEOM, 32, 2, 2, 2, 1>(p, max_n, st);
}

}  // namespace

BX_API int bx_conv_tc_ntile(int Cout) { return Cout > 64 ? 128 : (Cout > 32 ? 64 : 32); }

// Tuning knob (experiments / tests): maximum number of 16-channel stages accumulated in tensor memory
// before the accumulators are drained with a rounded fp32 add.  <= 0 restores the default.
static int g_tc_max_stages = 6;
BX_API int bx_conv_tc_set_segment_stages(int stages) {
    const int old = g_tc_max_stages;
    g_tc_max_stages = stages > 0 ? stages : 6;
    return old;
}

BX_API int bx_conv_layer_tc(int geom, const float *in, const float *w_tc, const float *bias, float *out, int n,
                            const int32_t *d_n, int Cin, int Cout, int D, int H, int W, int kd, int kh, int kw, int relu,
                            const float *equi_s, const float *equi_t, const int32_t *s_mids, const int32_t *t_mids,
                            void *stream) {
    BX_REQUIRE(w_tc && bias && out, "bx_conv_layer_tc: null pointer");
    BX_REQUIRE(n >= 0 && Cin >= 16 && Cin % 16 == 0 && Cout >= 4 && Cout % 4 == 0 && Cout <= 128, "bx_conv_layer_tc: bad channels Cin=%d Cout=%d", Cin, Cout);
    BX_REQUIRE((reinterpret_cast<uintptr_t>(w_tc) & 15) == 0, "bx_conv_layer_tc: weights must be 16-byte aligned");
    BX_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
               "bx_conv_layer_tc: activations and bias must be 16-byte aligned");
    BX_REQUIRE(kd >= 1 && kh >= 1 && kw >= 1 && (long long)kd * kh * kw <= TC_MAX_TAPS, "bx_conv_layer_tc: at most %d kernel taps", TC_MAX_TAPS);
    ConvTcParams p = {};
    p.in = in; p.w = w_tc; p.bias = bias; p.out = out; p.n = n; p.d_n = d_n;
    p.Cin = Cin; p.Cout = Cout; p.D = D; p.H = H; p.W = W; p.kd = kd; p.kh = kh; p.kw = kw; p.relu = relu;
    p.equi_s = equi_s; p.equi_t = equi_t; p.s_mids = s_mids; p.t_mids = t_mids;
    p.T = kd * kh * kw;
    p.seg_len = g_tc_max_stages;   // main-accumulator segment: 6 stages = 12 truncating accumulations
#ifdef BX_TC_TRACE
    static long long *d_trace = nullptr;
    const char *te = getenv("BX_TC_TRACE");
    const bool do_trace = te && atoi(te) == Cin * 1000 + Cout;
    if (do_trace) {
        if (!d_trace) cudaMalloc(&d_trace, sizeof(long long) * 4096);
        cudaMemset(d_trace, 0, sizeof(long long) * 4096);
        p.trace = d_trace;
    }
#endif
    cudaStream_t st = bx_stream(stream);
    switch (geom) {
        case BX_GEOM_CYL3D:
            BX_REQUIRE(in && D == 3 && H == 7 && W == 20 && kd == 3 && kh == 3 && kw == 3, "bx_conv_layer_tc: CYL3D expects [C,3,7,20], k=3x3x3");
            p.S_in = 420; p.S_out = 140; p.OD = 1; p.OH = 7; p.OW = 20;
            return dispatch_nt<BX_GEOM_CYL3D>(p, n, st);
        case BX_GEOM_CYL2D:
            BX_REQUIRE(in && D == 1 && H == 7 && W == 20 && kd == 1 && kh == 3 && kw == 3, "bx_conv_layer_tc: CYL2D expects [C,7,20], k=3x3");
            p.S_in = 140; p.S_out = 140; p.OD = 1; p.OH = 7; p.OW = 20;
#ifdef BX_TC_TRACE
            {   // debugging aid: BX_TC_TRACE=<Cin*1000+Cout> prints the stage timeline of the middle CTA of that layer once
                const int rc = dispatch_nt<BX_GEOM_CYL2D>(p, n, st);
                static int printed = 0;
                if (do_trace && printed < 1) {
                    ++printed;
                    cudaDeviceSynchronize();
                    static long long h[4096];
                    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
                    const long long t00 = h[0];
                    for (int g = 0; g < 4; ++g)
                        for (int k = 0; k < 24; ++k) {
                            const long long *r = h + (g * 64 + k) * 3;
                            if (r[2]) printf("L g%d k%2d top %7lld  empty-wait %5lld  fill %5lld\n", g, k, r[0] - t00, r[1] - r[0], r[2] - r[1]);
                        }
                    printf("CTA start %lld  end %lld  (total %lld cycles); a loader group left its main loop at %lld\n", h[4000] - t00, h[4001] - t00,
                           h[4001] - h[4000], h[4002] - t00);
                    const long long *m = h + 4 * 3 * 64;
                    for (int it = 0; it < 72 && m[it * 3]; ++it)
                        printf("M it%3d top %7lld (+%5lld)  acc/b-wait %5lld  a-wait %5lld\n", it, m[it * 3] - t00, it ? m[it * 3] - m[it * 3 - 3] : 0,
                               m[it * 3 + 1] - m[it * 3], m[it * 3 + 2] - m[it * 3 + 1]);
                }
                return rc;
            }
#else
            return dispatch_nt<BX_GEOM_CYL2D>(p, n, st);
#endif
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
        case BX_GEOM_COSTAB:
            BX_REQUIRE(equi_s && equi_t, "bx_conv_layer_tc: COSTAB needs the A and B factors");
            BX_REQUIRE(Cin == 32 && D == 18 && H == 3 && W == 18 && kd == 3 && kh == 3 && kw == 3, "bx_conv_layer_tc: COSTAB expects the [32,18,3,18] activation, k=3x3x3");
            p.OD = 16; p.OH = 1; p.OW = 16; p.S_in = 972; p.S_out = 256;
            return dispatch_nt<BX_GEOM_COSTAB>(p, n, st);
        default:
            bx_set_error("bx_conv_layer_tc: unknown geometry %d", geom);
            return BX_ERR_INVALID_ARG;
    }
}
