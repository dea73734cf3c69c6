// adaptive_weight.cu — a13: the "ConvNet" local aggregation of the TF tree (AdaptiveWeight with
// local_input_feature 'dp', shared_channels 1, fc_num 1, reduction 'mean', no softmax — the shipped
// config/s3dis/adapt.yaml:19-26; reference tensorflow/models/local_aggregation_operators.py:316-500):
//     dp[n,k]   = (support[idx[n,k]] - query[n]) / radius                       (:379-382)
//     w[n,k,c]  = W[c,:] . dp[n,k] + b[c]                                        (:426-430, one FC on dp)
//     out[n,c]  = sum_k w[n,k,c] * feat[idx[n,k], c] / (cnt[n] + 1e-5)           (:456-471)
// with the shadow neighbour (idx == n0) contributing a zero feature row (:370-372) and
// cnt[n] = #{k : idx[n,k] < max(idx)} (the reference's own "valid" test, :466-470).
// The reference materialises (n,K,c) three times (two tf.gather + the FC output); here a warp owns a
// point and a 128-channel slab and streams the K neighbour rows once.  Backward: d feat by scatter-add,
// dW/db by per-lane accumulation.
#include "common.cuh"

#define AW_THREADS 256
#define AW_CPL 4          // channels per lane -> 128-channel slab per warp task

__global__ void k_aw_maxidx(const int *__restrict__ idx, long long total, int *out)
{
    int m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        m = max(m, __ldg(idx + i));
    m = __reduce_max_sync(CB_FULL_MASK, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// MODE 0 forward, 1 backward
template <int MODE>
__global__ void __launch_bounds__(AW_THREADS) k_aw(int n, int k, int c, int n0, const float *__restrict__ qpts,
                                                   const float *__restrict__ spts, const int *__restrict__ idx,
                                                   const float *__restrict__ feat, const float *__restrict__ W,
                                                   const float *__restrict__ bias, float radius,
                                                   const int *__restrict__ pad_num, float *__restrict__ out,
                                                   const float *__restrict__ gout, float *__restrict__ gfeat,
                                                   float *__restrict__ gW, float *__restrict__ gb)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ch0 = blockIdx.y * 32 * AW_CPL;
    float w0[AW_CPL], w1[AW_CPL], w2[AW_CPL], bb[AW_CPL];
    float aW[AW_CPL][3], ab[AW_CPL];
    bool live[AW_CPL];
#pragma unroll
    for (int i = 0; i < AW_CPL; i++) {
        const int ch = ch0 + lane + 32 * i;
        live[i] = ch < c;
        w0[i] = live[i] ? W[3 * ch] : 0.f; w1[i] = live[i] ? W[3 * ch + 1] : 0.f; w2[i] = live[i] ? W[3 * ch + 2] : 0.f;
        bb[i] = live[i] ? bias[ch] : 0.f;
        aW[i][0] = aW[i][1] = aW[i][2] = 0.f; ab[i] = 0.f;
    }
    const int pad = __ldg(pad_num);
    const int warps = gridDim.x * (AW_THREADS / 32);
    const float inv_r = 1.0f;   // (the division by radius is kept per element, as the reference: dp / radius)
    (void)inv_r;
    for (int pt = blockIdx.x * (AW_THREADS / 32) + wib; pt < n; pt += warps) {
        const float qx = __ldg(qpts + 3 * pt), qy = __ldg(qpts + 3 * pt + 1), qz = __ldg(qpts + 3 * pt + 2);
        float acc[AW_CPL], g[AW_CPL];
#pragma unroll
        for (int i = 0; i < AW_CPL; i++) acc[i] = 0.f;
        // count of "valid" neighbours as the reference counts them (idx < max idx)
        int cnt = 0;
        for (int k0 = 0; k0 < k; k0 += 32) {
            const int kk = k0 + lane;
            cnt += __popc(__ballot_sync(CB_FULL_MASK, kk < k && __ldg(idx + (size_t)pt * k + kk) < pad));
        }
        const float inv = 1.0f / ((float)cnt + 1e-5f);
#pragma unroll
        for (int i = 0; i < AW_CPL; i++) g[i] = (MODE == 1 && live[i]) ? __ldg(gout + (size_t)pt * c + ch0 + lane + 32 * i) * inv : 0.f;
        // neighbours in chunks of 32: lane kk fetches ITS neighbour's index and relative position (one parallel gather
        // instead of K serial broadcast loads), then the chunk is walked 4 neighbours at a time with all feature loads
        // of the 4 rows in flight together
        for (int k0 = 0; k0 < k; k0 += 32) {
            const int kk = k0 + lane;
            int jl = n0;
            float dxl = 0.f, dyl = 0.f, dzl = 0.f;
            if (kk < k) {
                jl = __ldg(idx + (size_t)pt * k + kk);
                if (jl >= 0 && jl < n0) {
                    dxl = (__ldg(spts + 3 * jl) - qx) / radius; dyl = (__ldg(spts + 3 * jl + 1) - qy) / radius;
                    dzl = (__ldg(spts + 3 * jl + 2) - qz) / radius;
                } else {
                    jl = n0;
                }
            }
            const int kend = min(32, k - k0);
            for (int u0 = 0; u0 < kend; u0 += 4) {
                int j[4];
                float dx[4], dy[4], dz[4], f[4][AW_CPL];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int src = min(u0 + u, 31);
                    j[u] = __shfl_sync(CB_FULL_MASK, jl, src);
                    dx[u] = __shfl_sync(CB_FULL_MASK, dxl, src); dy[u] = __shfl_sync(CB_FULL_MASK, dyl, src);
                    dz[u] = __shfl_sync(CB_FULL_MASK, dzl, src);
                    if (u0 + u >= kend) j[u] = n0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int i = 0; i < AW_CPL; i++)
                        f[u][i] = (j[u] < n0 && live[i]) ? __ldg(feat + (size_t)j[u] * c + ch0 + lane + 32 * i) : 0.f;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (j[u] >= n0) continue;                          // shadow neighbour: zero feature row (warp-uniform)
#pragma unroll
                    for (int i = 0; i < AW_CPL; i++) {
                        if (!live[i]) continue;
                        const float w = w0[i] * dx[u] + w1[i] * dy[u] + w2[i] * dz[u] + bb[i];
                        if (MODE == 0) {
                            acc[i] += w * f[u][i];
                        } else {
                            atomicAdd(gfeat + (size_t)j[u] * c + ch0 + lane + 32 * i, g[i] * w);
                            const float gf = g[i] * f[u][i];
                            aW[i][0] += gf * dx[u]; aW[i][1] += gf * dy[u]; aW[i][2] += gf * dz[u]; ab[i] += gf;
                        }
                    }
                }
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < AW_CPL; i++)
                if (live[i]) out[(size_t)pt * c + ch0 + lane + 32 * i] = acc[i] * inv;
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < AW_CPL; i++) {
            if (!live[i]) continue;
            const int ch = ch0 + lane + 32 * i;
            atomicAdd(gW + 3 * ch, aW[i][0]); atomicAdd(gW + 3 * ch + 1, aW[i][1]); atomicAdd(gW + 3 * ch + 2, aW[i][2]);
            atomicAdd(gb + ch, ab[i]);
        }
    }
}

static dim3 aw_grid(int n, int c)
{
    int gx = (n + AW_THREADS / 32 - 1) / (AW_THREADS / 32);
    if (gx > 148 * 8) gx = 148 * 8;
    if (gx < 1) gx = 1;
    return dim3(gx, (c + 32 * AW_CPL - 1) / (32 * AW_CPL));
}

extern "C" int cb_adaptive_weight_forward(int n, int k, int c, int n0, const float *query_pts, const float *support_pts,
                                          const int *idx, const float *feat, const float *W, const float *bias,
                                          float radius, int *pad_num, float *out, void *stream)
{
    CB_REQUIRE(n >= 0 && k > 0 && c > 0 && n0 >= 0 && query_pts && support_pts && idx && feat && W && bias && pad_num && out &&
                   radius > 0.f, CB_EINVAL, "cb_adaptive_weight_forward: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(pad_num, 0, sizeof(int), st);
    if (n == 0) return CB_OK;
    const long long total = (long long)n * k;
    int g = (int)((total + 255) / 256);
    if (g > 148 * 4) g = 148 * 4;
    k_aw_maxidx<<<g, 256, 0, st>>>(idx, total, pad_num);                 // padding_num = reduce_max(neighbors_indices) (:466)
    k_aw<0><<<aw_grid(n, c), AW_THREADS, 0, st>>>(n, k, c, n0, query_pts, support_pts, idx, feat, W, bias, radius, pad_num, out,
                                                  nullptr, nullptr, nullptr, nullptr);
    CB_COUNT(2);
    CB_CUDA_CHECK("cb_adaptive_weight_forward");
    return CB_OK;
}

extern "C" int cb_adaptive_weight_backward(int n, int k, int c, int n0, const float *query_pts, const float *support_pts,
                                           const int *idx, const float *feat, const float *W, const float *bias,
                                           float radius, const int *pad_num, const float *grad_out, float *grad_feat,
                                           float *grad_W, float *grad_b, void *stream)
{
    CB_REQUIRE(n >= 0 && k > 0 && c > 0 && query_pts && support_pts && idx && feat && W && bias && pad_num && grad_out && grad_feat &&
                   grad_W && grad_b, CB_EINVAL, "cb_adaptive_weight_backward: bad arguments");
    if (n == 0) return CB_OK;
    k_aw<1><<<aw_grid(n, c), AW_THREADS, 0, (cudaStream_t)stream>>>(n, k, c, n0, query_pts, support_pts, idx, feat, W, bias, radius,
                                                                    pad_num, nullptr, grad_out, grad_feat, grad_W, grad_b);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_adaptive_weight_backward");
    return CB_OK;
}
