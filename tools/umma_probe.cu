// umma_probe.cu -- micro-benchmark behind DESIGN.md section 5: what paces the shifted-descriptor convolution kernel?
//   (1) SS-form tcgen05.mma rate as a function of N (shared-memory operand bytes per MMA: A 4 KB + B N*32 B) for
//       cta_group::1 (M = 128) and cta_group::2 (M = 256, every CTA supplies its 128 A rows and HALF of B);
//   (2) tcgen05.ld drain rate (128 lanes x C columns) alone and while the MMAs run;
//   (3) the same with a bulk-copy stream into shared memory beside it.
// Stand-alone:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/umma_probe tools/umma_probe.cu
//               tools/_bin/umma_probe            (prints one line per configuration; all 148 SMs busy)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../buffer-x_b200/csrc/bx_tcgen05.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int AROWS = 176, KCORE = AROWS * 16, CHUNK = 4 * KCORE;     // the conv_sd A chunk image [split][kcore][row][16 B]

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__device__ __forceinline__ void mma_f16(uint32_t leader, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %0, 0;\n\tmov.b64 da, {%2, %4};\n\tmov.b64 db, {%3, %4};\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %5, p;\n\t}\n" ::"r"(leader), "r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %0, 0;\n\tmov.b64 da, {%2, %4};\n\tmov.b64 db, {%3, %4};\n\t"
                     "@q tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %5, p;\n\t}\n" ::"r"(leader), "r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t leader, uint32_t bar) {
    if constexpr (CG == 1) mma_commit(leader, bar);
    else
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%1], %2;\n\t}\n" ::"r"(leader),
                     "r"(bar), "h"((unsigned short)3) : "memory");
}


// variants of the cta_group::1 instruction (probe only)
__device__ __forceinline__ void mma_f16_hint(uint32_t leader, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc, int last) {
    if (last)
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %0, 0;\n\tmov.b64 da, {%2, %4};\n\tmov.b64 db, {%3, %4};\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%1], da, db, %5, p;\n\t}\n" ::"r"(leader), "r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %0, 0;\n\tmov.b64 da, {%2, %4};\n\tmov.b64 db, {%3, %4};\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%1], da, db, %5, p;\n\t}\n" ::"r"(leader), "r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t leader, uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %0, 0;\n\tmov.b64 db, {%3, %4};\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], db, %5, p;\n\t}\n" ::"r"(leader), "r"(d), "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ws(uint32_t leader, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc, int use) {
    if (use)
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %0, 0;\n\tmov.b64 da, {%2, %4};\n\tmov.b64 db, {%3, %4};\n\t"
                     "@q tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%1], da, db, %5, p;\n\t}\n" ::"r"(leader), "r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %0, 0;\n\tmov.b64 da, {%2, %4};\n\tmov.b64 db, {%3, %4};\n\t"
                     "@q tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%1], da, db, %5, p;\n\t}\n" ::"r"(leader), "r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}

struct Args {
    int N;            // MMA N (full N of the instruction)
    int iters;        // outer iterations
    int nmma;         // MMAs per iteration (9 taps x variants)
    int do_mma;       // 0: no MMAs
    int drain_cols;   // columns drained per iteration by the drain warps (0: no drains)
    int drain_warps;  // 4 or 8
    int copy_bytes;   // bytes bulk-copied into shared memory per iteration (0: none)
    int variant;      // 0 two accumulators, 1 four, 2 MMA pairs share A, 3 same + collector::a hints, 4 A from tensor memory, 5 .ws with B pairs shared (collector::b0), 6 M = 64
    const unsigned char *src;
    unsigned long long *out;
};

template <int CG, int V, int NN>
__global__ void __launch_bounds__(10 * 32, 1) probe(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[8];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
    if (warp == 8) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 2 * CHUNK, c_base = a_base + 120 * 1024, bar_base = smem_u32(&bars[0]);
    unsigned long long t0 = 0, t1 = 0;

    if (warp == 8) {
        if (a.do_mma && rank == 0) {
            constexpr int N = NN, nb = CG == 2 ? N / 2 : N;              // B rows held by one CTA
            constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24);
            const uint32_t DESC_HI = (128u >> 4) | (1u << 14);
            const uint32_t a0 = (a_base >> 4) | (((uint32_t)KCORE >> 4) << 16);
            const uint32_t b0 = (b_base >> 4) | (((uint32_t)(nb * 16) >> 4) << 16);
            constexpr uint32_t bstage = (uint32_t)(nb * 32) >> 4;             // one tap of [kcore][n][16 B]
            const uint32_t leader = elect_leader();
            t0 = clock64();
            for (int it = 0; it < a.iters; ++it) {
#pragma unroll
                for (int j = 0; j < 18; ++j) {
                    const int tap = j % 9;
                    const uint32_t ad = a0 + (uint32_t)((it & 1) * (CHUNK >> 4)) + (uint32_t)((tap / 3) * 22 + tap % 3) + (uint32_t)((j / 9) & 1) * ((2u * KCORE) >> 4);
                    const uint32_t bd = b0 + (uint32_t)tap * bstage;
                    if constexpr (CG == 2 || V == 0) mma_f16<CG>(leader, tmem_base + (uint32_t)((j & 1) * (N >= 256 ? 256 : N)), ad, bd, DESC_HI, IDESC, j >= 2 ? 1u : 0u);
                    else if constexpr (V == 1) mma_f16<1>(leader, tmem_base + (uint32_t)((j & 3) * N), ad, bd, DESC_HI, IDESC, j >= 4 ? 1u : 0u);
                    else if constexpr (V == 2 || V == 3) {
                        const int t2 = (j >> 1) % 9;
                        const uint32_t ad2 = a0 + (uint32_t)((it & 1) * (CHUNK >> 4)) + (uint32_t)((t2 / 3) * 22 + t2 % 3);
                        if constexpr (V == 2) mma_f16<1>(leader, tmem_base + (uint32_t)((j & 1) * N), ad2, bd, DESC_HI, IDESC, j >= 2 ? 1u : 0u);
                        else mma_f16_hint(leader, tmem_base + (uint32_t)((j & 1) * N), ad2, bd, DESC_HI, IDESC, j >= 2 ? 1u : 0u, j & 1);
                    } else if constexpr (V == 4) mma_f16_ts(leader, tmem_base + (uint32_t)((j & 1) * N), tmem_base + 384u + (uint32_t)(tap * 8), bd, DESC_HI, IDESC, j >= 2 ? 1u : 0u);
                    else if constexpr (V == 5) {
                        const int t2 = (j >> 1) % 9;
                        mma_f16_ws(leader, tmem_base + (uint32_t)((j & 1) * N), ad, b0 + (uint32_t)t2 * bstage, DESC_HI, IDESC, j >= 2 ? 1u : 0u, j & 1);
                    } else if constexpr (V == 6) {
                        const uint32_t ID64 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
                        mma_f16<1>(leader, tmem_base + (uint32_t)((j & 1) * N), ad, bd, DESC_HI, ID64, j >= 2 ? 1u : 0u);
                    }
                }
                commit<CG>(leader, bar_base + 8u * (uint32_t)(it & 1));
                if (it >= 1) mbar_wait(bar_base + 8u * (uint32_t)((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));   // at most two iterations in flight
            }
            mbar_wait(bar_base + 8u * (uint32_t)((a.iters - 1) & 1), (uint32_t)(((a.iters - 1) >> 1) & 1));
            t1 = clock64();
            if (lane == 0) { a.out[blockIdx.x * 4 + 0] = t1 - t0; }
        } else if (lane == 0) a.out[blockIdx.x * 4 + 0] = 0;
    } else if (warp == 9) {
        if (a.copy_bytes && lane == 0) {
            // free-running bulk-copy stream into a scratch region (two buffers)
            t0 = clock64();
            for (int it = 0; it < a.iters; ++it) {
                const uint32_t b = 2u + (uint32_t)(it & 1);
                if (it >= 2) mbar_wait(bar_base + 8u * b, (uint32_t)(((it >> 1) - 1) & 1));
                mbar_arrive_expect_tx(bar_base + 8u * b, (uint32_t)a.copy_bytes);
                bulk_g2s(c_base + (uint32_t)(it & 1) * 40960u, a.src + (size_t)((it * 148 + blockIdx.x) % 2048) * 40960, (uint32_t)a.copy_bytes, bar_base + 8u * b);
            }
            for (int it = a.iters > 2 ? a.iters - 2 : 0; it < a.iters; ++it) mbar_wait(bar_base + 8u * (2u + (uint32_t)(it & 1)), (uint32_t)((it >> 1) & 1));
            t1 = clock64();
            a.out[blockIdx.x * 4 + 2] = t1 - t0;
        } else if (lane == 0) a.out[blockIdx.x * 4 + 2] = 0;
    } else if (warp < a.drain_warps && a.drain_cols > 0) {
        const uint32_t tm = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * (a.drain_cols / (a.drain_warps / 4)));
        const int cols = a.drain_cols / (a.drain_warps / 4);
        float acc = 0.f;
        t0 = clock64();
        for (int it = 0; it < a.iters; ++it) {
            for (int c0 = 0; c0 < cols; c0 += 32) {
                uint32_t v[32];
                tmem_ld<32>(tm + (uint32_t)c0, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e = 0; e < 32; ++e) acc += __uint_as_float(v[e]);
            }
        }
        t1 = clock64();
        if (lane == 0 && warp == 0) a.out[blockIdx.x * 4 + 1] = t1 - t0;
        if (acc == 123.456f) a.out[0] = 0;
    } else if (warp == 0 && lane == 0) a.out[blockIdx.x * 4 + 1] = 0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    if (warp == 8) {
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

static unsigned long long *d_out;
static unsigned char *d_src;

template <int CG, int V = 0, int NN = 128>
static void run(const char *name, Args a) {
    a.N = NN; a.nmma = 18; a.variant = V;
    a.out = d_out; a.src = d_src;
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(probe<CG, V, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaMemset(d_out, 0, 148 * 4 * 8));
    for (int rep = 0; rep < 2; ++rep) { CK(cudaLaunchKernelEx(&cfg, probe<CG, V, NN>, a)); CK(cudaDeviceSynchronize()); }
    std::vector<unsigned long long> h(148 * 4);
    CK(cudaMemcpy(h.data(), d_out, 148 * 4 * 8, cudaMemcpyDeviceToHost));
    double mx[3] = {0, 0, 0};
    for (int b = 0; b < 148; ++b) for (int k = 0; k < 3; ++k) if ((double)h[b * 4 + k] > mx[k]) mx[k] = (double)h[b * 4 + k];
    const double per_mma = a.do_mma ? mx[0] / ((double)a.iters * a.nmma) : 0.0;
    const double floor_cyc = a.N / 2.0;       // M = 128 (256 for the pair): N/2 cycles per K = 16 MMA
    const int nb = CG == 2 ? a.N / 2 : a.N;
    printf("%-34s cg%d N=%3d  mma %7.1f cyc (floor %5.1f, %5.1f %%)  smem operand %5.1f B/clk/SM | drain %5d cols x %d warps: %8.1f cyc/iter (%6.1f B/clk) | copy %6d B/iter: %8.1f cyc/iter (%5.1f B/clk)\n",
           name, CG, a.N, per_mma, floor_cyc, per_mma > 0 ? 100.0 * floor_cyc / per_mma : 0.0, per_mma > 0 ? (4096.0 + nb * 32.0) / per_mma : 0.0, a.drain_cols, a.drain_warps,
           mx[1] / a.iters, mx[1] > 0 ? 128.0 * a.drain_cols * 4.0 / (mx[1] / a.iters) : 0.0, a.copy_bytes, mx[2] / a.iters, mx[2] > 0 ? a.copy_bytes / (mx[2] / a.iters) : 0.0);
}

#define RUN3(CG, V, NAME, ...) { Args a = __VA_ARGS__; run<CG, V, 64>(NAME, a); run<CG, V, 128>(NAME, a); run<CG, V, 256>(NAME, a); }
int main() {
    CK(cudaMalloc(&d_out, 148 * 4 * 8));
    CK(cudaMalloc(&d_src, (size_t)2048 * 40960));
    CK(cudaMemset(d_src, 0, (size_t)2048 * 40960));
    const int IT = 400;
    RUN3(1, 0, "mma only", {0, IT, 18, 1, 0, 4, 0});
    { Args a = {0, IT, 18, 1, 0, 4, 0}; run<1, 0, 32>("mma only", a); }
    RUN3(2, 0, "mma only (pair)", {0, IT, 18, 1, 0, 4, 0});
    for (int w : {4, 8}) for (int c : {128, 256}) { Args a = {128, IT, 18, 0, c, w, 0}; run<1, 0, 128>("drain only", a); }
    RUN3(1, 0, "mma + drain 128 cols", {0, IT, 18, 1, 128, 8, 0});
    RUN3(1, 0, "mma + 40 KB copy / iter", {0, IT, 18, 1, 0, 4, 40960});
    RUN3(1, 0, "mma + drain + copy", {0, IT, 18, 1, 128, 8, 40960});
    RUN3(2, 0, "pair: mma + drain + 20 KB copy", {0, IT, 18, 1, 128, 8, 20480});
    { Args a = {128, IT, 18, 0, 0, 4, 40960}; run<1, 0, 128>("copy only", a); }
    { Args a = {0, IT, 18, 1, 0, 4, 0}; run<1, 1, 32>("v1 four accumulators", a); run<1, 1, 64>("v1 four accumulators", a); run<1, 1, 128>("v1 four accumulators", a); }
    RUN3(1, 2, "v2 MMA pairs share A", {0, IT, 18, 1, 0, 4, 0});
    RUN3(1, 3, "v3 pairs share A + collector::a", {0, IT, 18, 1, 0, 4, 0});
    { Args a = {0, IT, 18, 1, 0, 4, 0}; run<1, 4, 32>("v4 A from tensor memory", a); run<1, 4, 64>("v4 A from tensor memory", a); run<1, 4, 128>("v4 A from tensor memory", a); }
    RUN3(1, 6, "v6 M = 64", {0, IT, 18, 1, 0, 4, 0});
    RUN3(1, 5, "v5 .ws, pairs share B (b0)", {0, IT, 18, 1, 0, 4, 0});
    return 0;
}
