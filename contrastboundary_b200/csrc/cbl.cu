// cbl.cu — a9: fused contrastive-boundary loss of one stage (forward + backward).
// Reference: ContrastHead.point_contrast + dist_l2 + posmask_cnt + contrast_softnn
// (pytorch/model/heads.py:185-246,116-119,145-165) with sub-scene labels from
// basic_operators.py:9-50.  Per point i with neighbours idx[i,1:] (column 0 = the point itself):
//     pos_k  = [argmax label_i == argmax label_{nbr k}]            boundary point iff 0 < sum pos < K-1
//     dist_k = sqrt(|f_i - f_nbr|^2 + 1e-12)
//     e_k    = exp((-dist_k - max_k(-dist_k)) / T)
//     loss_i = -log(sum_k e_k pos_k / sum_k e_k + 1e-12)            stage loss = w * mean over boundary points
// The reference gathers (m,K-1,ncls) and (m,K-1,d) tensors and runs ~12 kernels plus a host sync
// (torch.any, heads.py:222); here one warp owns a point, LANE = NEIGHBOUR: each lane streams its
// neighbour's d-float feature row with LDG.128 and keeps the whole soft-NN in registers.
#include "common.cuh"

#define CBL_THREADS 256
#define CBL_EPS 1e-12f

// class of every point: level 0 -> target; deeper levels -> arg-max of the histogram of the kr nearest
// full-resolution labels (first maximum, as torch.argmax on the mean one-hot).
__global__ void k_cbl_classes(int m, int kr, int ncls, const int *__restrict__ label_idx,
                              const long long *__restrict__ target, int *__restrict__ cls)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (!label_idx) { cls[i] = (int)target[i]; return; }
    int cnt[64];
    for (int c = 0; c < ncls; c++) cnt[c] = 0;
    for (int t = 0; t < kr; t++) {
        const int l = (int)target[__ldg(label_idx + (size_t)i * kr + t)];
        if (l >= 0 && l < ncls) cnt[l]++;
    }
    int best = 0;
    for (int c = 1; c < ncls; c++) if (cnt[c] > cnt[best]) best = c;
    cls[i] = best;
}

template <int D>
__device__ __forceinline__ void cbl_load_row(const float *__restrict__ p, float (&v)[D])
{
#pragma unroll
    for (int c = 0; c < D; c += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p + c));
        v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
    }
}

// mode 0: forward (accumulate sums[0] += loss_i, sums[1] += 1 over boundary points)
// mode 1: backward (gfeat += d loss / d feat, scale = *scale_ptr per boundary point)
template <int D, int MODE>
__global__ void __launch_bounds__(CBL_THREADS) k_cbl(int m, int K, const float *__restrict__ feat,
                                                     const int *__restrict__ idx, const int *__restrict__ cls,
                                                     float inv_t, float *__restrict__ sums,
                                                     const float *__restrict__ scale_ptr, float *__restrict__ gfeat,
                                                     int n_valid, int flavour)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warps = gridDim.x * (CBL_THREADS / 32);
    const int nk = K - 1;
    float acc_loss = 0.f, acc_cnt = 0.f;
    const float scale = MODE == 1 ? __ldg(scale_ptr) : 0.f;
    for (int i = blockIdx.x * (CBL_THREADS / 32) + wib; i < m; i += warps) {
        const int ci = __ldg(cls + i);
        // pass 1: positives / negatives among VALID neighbours (cheap) -> skip non-boundary points early
        int npos = 0, nneg = 0;
        for (int k0 = 0; k0 < nk; k0 += 32) {
            const int k = k0 + lane;
            bool pos = false, ok = false;
            if (k < nk) {
                const int j = __ldg(idx + (size_t)i * K + 1 + k);
                ok = j < n_valid;                                    // shadow neighbours (TF radius search) are invalid
                if (ok) pos = __ldg(cls + j) == ci;
            }
            npos += __popc(__ballot_sync(CB_FULL_MASK, pos));
            nneg += __popc(__ballot_sync(CB_FULL_MASK, ok && !pos));
        }
        if (!(npos > 0 && nneg > 0)) continue;                      // heads.py:213-214 / head.py:641-662
        float f[D];
        cbl_load_row<D>(feat + (size_t)i * D, f);
        // per-lane neighbour state for up to 2 rounds (K-1 <= 64)
        float dist[2], e[2];
        bool posm[2], val[2], clampd[2] = {false, false};
        int nb[2];
        float mx = -3.0e38f;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int k = r * 32 + lane;
            val[r] = k < nk;
            nb[r] = val[r] ? __ldg(idx + (size_t)i * K + 1 + k) : i;
            if (nb[r] >= n_valid) { val[r] = false; nb[r] = i; }
            posm[r] = val[r] && (__ldg(cls + nb[r]) == ci);
            float g[D];
            cbl_load_row<D>(feat + (size_t)nb[r] * D, g);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < D; c++) { const float t = f[c] - g[c]; s = fmaf(t, t, s); }
            dist[r] = flavour == 0 ? sqrtf(s + CBL_EPS) : sqrtf(fmaxf(s, CBL_EPS));   // heads.py:116-119 / head.py:183-185
            if (flavour == 1 && s < CBL_EPS) clampd[r] = true;
            if (val[r]) mx = fmaxf(mx, -dist[r]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(CB_FULL_MASK, mx, o));
        float pos = 0.f, neg = 0.f;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            e[r] = val[r] ? expf((-dist[r] - mx) * inv_t) : 0.f;
            neg += e[r];
            if (posm[r]) pos += e[r];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pos += __shfl_xor_sync(CB_FULL_MASK, pos, o);
            neg += __shfl_xor_sync(CB_FULL_MASK, neg, o);
        }
        const float ratio = pos / neg;
        if (MODE == 0) {
            if (lane == 0) { acc_loss += -logf(ratio + CBL_EPS); acc_cnt += 1.f; }
        } else {
            // dL/d dist_k = scale * (-1/(ratio+eps)) * e_k (pos_k*neg - pos)/neg^2 * (-inv_t)
            const float outer = scale * (-1.0f / (ratio + CBL_EPS)) / (neg * neg) * (-inv_t);
            float gi[D];
#pragma unroll
            for (int c = 0; c < D; c++) gi[c] = 0.f;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (!val[r]) continue;
                const float ddist = outer * e[r] * ((posm[r] ? neg : 0.f) - pos);
                const float coef = clampd[r] ? 0.f : ddist / dist[r];   // d dist / d s = 1/(2 dist); d s / d f = 2 (f - g); max() clamps -> 0
                float g[D];
                cbl_load_row<D>(feat + (size_t)nb[r] * D, g);
                float *dst = gfeat + (size_t)nb[r] * D;
#pragma unroll
                for (int c = 0; c < D; c += 4) {
                    const float t0 = coef * (f[c] - g[c]), t1 = coef * (f[c + 1] - g[c + 1]), t2 = coef * (f[c + 2] - g[c + 2]),
                                t3 = coef * (f[c + 3] - g[c + 3]);
                    gi[c] += t0; gi[c + 1] += t1; gi[c + 2] += t2; gi[c + 3] += t3;
                    atomicAdd(reinterpret_cast<float4 *>(dst + c), make_float4(-t0, -t1, -t2, -t3));
                }
            }
            // sum the per-lane contributions to d f_i: butterfly per channel, lane c%32 keeps channel c
#pragma unroll
            for (int c = 0; c < D; c++) {
                float v = gi[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CB_FULL_MASK, v, o);
                if (lane == (c & 31)) atomicAdd(gfeat + (size_t)i * D + c, v);
            }
        }
    }
    if (MODE == 0) {
        __shared__ float red[2][CBL_THREADS / 32];
        if (lane == 0) { red[0][wib] = acc_loss; red[1][wib] = acc_cnt; }
        __syncthreads();
        if (threadIdx.x < 2) {
            float t = 0.f;
            for (int w = 0; w < CBL_THREADS / 32; w++) t += red[threadIdx.x][w];
            if (t != 0.f) atomicAdd(sums + threadIdx.x, t);
        }
    }
}

extern "C" int cb_cbl_classes(int m, int kr, int ncls, const int *label_idx, const long long *target, int *cls, void *stream)
{
    CB_REQUIRE(m >= 0 && ncls > 0 && ncls <= 64 && target && cls, CB_EINVAL, "cb_cbl_classes: bad arguments (ncls <= 64)");
    if (m == 0) return CB_OK;
    k_cbl_classes<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(m, kr, ncls, label_idx, target, cls);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_cbl_classes");
    return CB_OK;
}

template <int MODE>
static int cbl_launch(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                      float *sums, const float *scale, float *gfeat, int n_valid, int flavour, cudaStream_t st)
{
    int grid = (m + CBL_THREADS / 32 - 1) / (CBL_THREADS / 32);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    const float inv_t = 1.0f / temperature;
    switch (D) {
    case 32: k_cbl<32, MODE><<<grid, CBL_THREADS, 0, st>>>(m, K, feat, idx, cls, inv_t, sums, scale, gfeat, n_valid, flavour); break;
    case 64: k_cbl<64, MODE><<<grid, CBL_THREADS, 0, st>>>(m, K, feat, idx, cls, inv_t, sums, scale, gfeat, n_valid, flavour); break;
    case 72: k_cbl<72, MODE><<<grid, CBL_THREADS, 0, st>>>(m, K, feat, idx, cls, inv_t, sums, scale, gfeat, n_valid, flavour); break;
    default:
        cb_set_error("cb_cbl: feature dim %d unsupported (32, 64, 72)", D);
        return CB_EUNSUPPORTED;
    }
    return CB_OK;
}

extern "C" int cb_cbl_forward_ex(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                                 float *sums, int n_valid, int flavour, void *stream);
extern "C" int cb_cbl_backward_ex(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                                  const float *scale, float *grad_feat, int n_valid, int flavour, void *stream);
extern "C" int cb_cbl_forward(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                              float *sums, void *stream)
{
    return cb_cbl_forward_ex(m, K, D, feat, idx, cls, temperature, sums, 0x7fffffff, 0, stream);
}
extern "C" int cb_cbl_backward(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                               const float *scale, float *grad_feat, void *stream)
{
    return cb_cbl_backward_ex(m, K, D, feat, idx, cls, temperature, scale, grad_feat, 0x7fffffff, 0, stream);
}

extern "C" int cb_cbl_forward_ex(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                                 float *sums, int n_valid, int flavour, void *stream)
{
    CB_REQUIRE(m >= 0 && K >= 2 && K <= 65 && feat && idx && cls && sums && temperature > 0.f, CB_EINVAL,
               "cb_cbl_forward: bad arguments (2 <= K <= 65)");
    if (m == 0) return CB_OK;
    int rc = cbl_launch<0>(m, K, D, feat, idx, cls, temperature, sums, nullptr, nullptr, n_valid, flavour, (cudaStream_t)stream);
    if (rc) return rc;
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_cbl_forward");
    return CB_OK;
}

extern "C" int cb_cbl_backward_ex(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                                  const float *scale, float *grad_feat, int n_valid, int flavour, void *stream)
{
    CB_REQUIRE(m >= 0 && K >= 2 && K <= 65 && feat && idx && cls && scale && grad_feat && temperature > 0.f, CB_EINVAL,
               "cb_cbl_backward: bad arguments");
    if (m == 0) return CB_OK;
    int rc = cbl_launch<1>(m, K, D, feat, idx, cls, temperature, nullptr, scale, grad_feat, n_valid, flavour, (cudaStream_t)stream);
    if (rc) return rc;
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_cbl_backward");
    return CB_OK;
}


// ---------------------------------------------------------------------------------------------
// SURVEY 8(f) row 3 — test-time boundary / plain masks (pytorch/model/basic_operators.py:69-97, used by
// pytorch/tool/test.py:392-428 on full-resolution rooms with kr in {16, 32, 64}).
//   valid neighbour = neighbour label >= 0;  bound_cnt = #{valid neighbours with a different label};
//   bound = bound_cnt > 0;  plain = every neighbour is invalid or carries the centre's label;  both & valid_mask.
// One thread per point; the reference materialises the (n, kr) gathered-label matrix three times.
// ---------------------------------------------------------------------------------------------
__global__ void k_boundary_mask(long long n, int kr, const long long *__restrict__ labels, const int *__restrict__ idx,
                                const unsigned char *__restrict__ valid, int *__restrict__ bound_cnt,
                                unsigned char *__restrict__ bound, unsigned char *__restrict__ plain)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long li = labels[i];
    int cnt = 0;
    bool all_eq = true;
    for (int t = 0; t < kr; t++) {
        const long long lj = labels[__ldg(idx + i * kr + t)];
        const bool vn = lj >= 0;
        if (vn && lj != li) { cnt++; all_eq = false; }
    }
    const bool v = valid ? valid[i] != 0 : true;
    if (bound_cnt) bound_cnt[i] = v ? cnt : 0;
    if (bound) bound[i] = (v && cnt > 0) ? 1 : 0;
    if (plain) plain[i] = (v && all_eq) ? 1 : 0;
}

extern "C" int cb_boundary_mask(long long n, int kr, const long long *labels, const int *neighbor_idx,
                                const unsigned char *valid_mask, int *bound_cnt, unsigned char *bound, unsigned char *plain,
                                void *stream)
{
    CB_REQUIRE(n >= 0 && kr >= 1 && labels && neighbor_idx, CB_EINVAL, "cb_boundary_mask: bad arguments");
    if (n == 0) return CB_OK;
    k_boundary_mask<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, kr, labels, neighbor_idx, valid_mask,
                                                                                  bound_cnt, bound, plain);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_boundary_mask");
    return CB_OK;
}
