// bx_spt.cu -- a6+a7: spherical-voxel point transformer fused with the point layer and its max-pool.
//
// Replaces MiniSpinNet.SPT (/root/reference/models/patch_embedder.py:150-165) = get_voxel_coordinate +
// sphere_query + var_to_invar (/root/reference/utils/common.py:422-498) and the 1x1 conv + BN + ReLU +
// max over the 10 samples (patch_embedder.py:26-30, 73-77).  The reference writes [K,420,10] indices,
// [K,420,10,3] points (75.6 MB at K=1500) and [K,16,420,10] activations (403 MB) to HBM per call; here
// one CTA owns a patch: its P points sit in shared memory (SoA), each warp walks voxels, finds the first
// `nv` in-ball points in index order with ballot/popcount, de-rotates them, applies the folded 3->16
// affine map + ReLU, max-reduces, and the [16,V] tile leaves the SM once, coalesced (26.9 KB/patch).
//
// Bit contract for the integer selection (oracle bxo_spt): d2 = ((qx-x)^2+(qy-y)^2)+(qz-z)^2 < r*r,
// first nv hits in index order; slot 0 zeroed when its index is 0 (utils/common.py:447-449), padding
// slots zeroed.  De-rotation x' = x*c + y*(-s), y' = x*s + y*c, z' = z.  -fmad=false.
#include "bx_common.cuh"

namespace {

constexpr int SPT_WARPS = 8;
constexpr int MAX_NV = 16;

__global__ void __launch_bounds__(SPT_WARPS * 32)
spt_pnt_kernel(const float *__restrict__ delta, int K, int P, const float *__restrict__ voxels, int V, int azi_n,
               const float *__restrict__ rot, float voxel_r, int nv, const float *__restrict__ w,
               const float *__restrict__ b, float *__restrict__ feat, int *__restrict__ dbg_vidx,
               float *__restrict__ dbg_inv) {
    extern __shared__ float smem[];
    float *px = smem;            // P
    float *py = px + P;          // P
    float *pz = py + P;          // P
    float *ft = pz + P;          // 16*V
    float *vx = ft + 16 * V;     // 3*V
    float *sw = vx + 3 * V;      // 64 (w[16][3], b[16])
    float *srot = sw + 64;       // 2*azi_n
    int *sel = reinterpret_cast<int *>(srot + 2 * azi_n);  // SPT_WARPS * MAX_NV

    const int k = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *dl = delta + (size_t)k * P * 3;
    for (int i = tid; i < 3 * P; i += blockDim.x) {
        const float v = dl[i];
        const int s = i / 3, c = i - 3 * s;
        (c == 0 ? px : (c == 1 ? py : pz))[s] = v;
    }
    for (int i = tid; i < 3 * V; i += blockDim.x) vx[i] = voxels[i];
    if (tid < 48) sw[tid] = w[tid];
    if (tid < 16) sw[48 + tid] = b[tid];
    for (int i = tid; i < 2 * azi_n; i += blockDim.x) srot[i] = rot[i];
    __syncthreads();

    const float r2 = voxel_r * voxel_r;
    int *my = sel + warp * MAX_NV;
    const int ch = lane & 15, half = lane >> 4;
    const float w0 = sw[3 * ch], w1 = sw[3 * ch + 1], w2 = sw[3 * ch + 2], bb = sw[48 + ch];
    const float empty_val = fmaxf(bb, 0.0f);  // relu(bn(conv(0))) -- what a zeroed slot contributes

    for (int v = warp; v < V; v += SPT_WARPS) {
        const float qx = vx[3 * v], qy = vx[3 * v + 1], qz = vx[3 * v + 2];
        int cnt = 0;
        for (int base = 0; base < P; base += 32) {
            const int i = base + lane;
            bool hit = false;
            if (i < P) hit = bx_d2(qx - px[i], qy - py[i], qz - pz[i]) < r2;
            const unsigned m = __ballot_sync(BX_FULL, hit);
            if (m) {
                const int slot = cnt + __popc(m & ((1u << lane) - 1u));
                if (hit && slot < nv) my[slot] = i;
                cnt += __popc(m);
                if (cnt >= nv) break;
            }
        }
        if (cnt > nv) cnt = nv;
        __syncwarp();
        const int first = cnt > 0 ? my[0] : 0;
        const int a = v % azi_n;
        const float cs = srot[2 * a], sn = srot[2 * a + 1];
        // live slots: l < cnt, except slot 0 when its index is 0
        const int l0 = (first == 0) ? 1 : 0;
        const bool any_zero = (cnt < nv) || (first == 0);
        float best = any_zero ? empty_val : -INFINITY;
        for (int l = l0 + half; l < cnt; l += 2) {
            // two half-warps interleave the live slots; (l0+half) may skip slot parity, handled by stride 2
            const int i = my[l];
            const float x = px[i], y = py[i], z = pz[i];
            const float xr = (x * cs) + (y * (-sn));
            const float yr = (x * sn) + (y * cs);
            const float val = (((w0 * xr) + (w1 * yr)) + (w2 * z)) + bb;
            best = fmaxf(best, fmaxf(val, 0.0f));
        }
        best = fmaxf(best, __shfl_xor_sync(BX_FULL, best, 16));
        if (lane < 16) ft[ch * V + v] = best;
        if (dbg_vidx || dbg_inv) {
            if (lane < nv) {
                const int l = lane;
                const int i = (l < cnt) ? my[l] : first;
                const size_t o = ((size_t)k * V + v) * nv + l;
                if (dbg_vidx) dbg_vidx[o] = i;
                if (dbg_inv) {
                    const bool live = (l < cnt) && !(l == 0 && first == 0);
                    float xr = 0.f, yr = 0.f, zr = 0.f;
                    if (live) {
                        const float x = px[i], y = py[i], z = pz[i];
                        xr = (x * cs) + (y * (-sn));
                        yr = (x * sn) + (y * cs);
                        zr = z;
                    }
                    dbg_inv[3 * o] = xr; dbg_inv[3 * o + 1] = yr; dbg_inv[3 * o + 2] = zr;
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    float *out = feat + (size_t)k * 16 * V;
    for (int i = tid; i < 16 * V; i += blockDim.x) out[i] = ft[i];
}

}  // namespace

BX_API int bx_spt_pnt(const float *delta, int K, int P, const float *voxels, int V, int azi_n, const float *rot,
                      float voxel_r, int nv, const float *w, const float *b, float *feat, int32_t *dbg_vidx,
                      float *dbg_inv, void *stream) {
    BX_REQUIRE(delta && voxels && rot && w && b && feat, "bx_spt_pnt: null pointer");
    BX_REQUIRE(K >= 0 && P >= 1 && V >= 1 && azi_n >= 1 && nv >= 1 && nv <= MAX_NV, "bx_spt_pnt: bad sizes");
    if (K == 0) return BX_OK;
    const size_t smem = sizeof(float) * (3 * (size_t)P + 16 * (size_t)V + 3 * (size_t)V + 64 + 2 * (size_t)azi_n) +
                        sizeof(int) * SPT_WARPS * MAX_NV;
    BX_REQUIRE(smem <= 200 * 1024, "bx_spt_pnt: P=%d V=%d needs %zu bytes of shared memory", P, V, smem);
    static size_t attr = 0;
    if (smem > attr) {
        BX_CUDA(cudaFuncSetAttribute(spt_pnt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    spt_pnt_kernel<<<K, SPT_WARPS * 32, smem, bx_stream(stream)>>>(delta, K, P, voxels, V, azi_n, rot, voxel_r, nv, w, b,
                                                                  feat, dbg_vidx, dbg_inv);
    BX_LAUNCH_CHECK();
    return BX_OK;
}
