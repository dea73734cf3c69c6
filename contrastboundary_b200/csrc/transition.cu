// transition.cu — a5: fused TransitionDown (stride > 1), forward + backward.
// Reference: TransitionDown.forward, pytorch/model/blocks.py:69-73:
//     g = cat(p[idx] - p_new, x[idx])  (m,k,3+c) ; y = Linear(3+c -> c', no bias)(g) ; BN ; ReLU ; max over k
// The linear is applied BEFORE the gather:  y[m,k,:] = Wxyz (p[idx]-p_new) + (x Wf^T)[idx]  with W = [Wxyz | Wf]
// (4x fewer FLOPs: n rows instead of m*k = 4n rows), so the (m,k,3+c) and (m,k,c') tensors never exist:
//     T1  per-channel sum / sum-of-squares of y            (training-mode BatchNorm statistics)
//     T2  out[m,c'] = max_k relu(bn(y)),  arg-max k saved as uint8
//     T3  backward sums over the arg-max entries           (BatchNorm backward needs mean(dy), mean(dy*yhat))
//     T4  dense backward: dy = gamma*invstd*(dy_sparse - mean(dy) - yhat*mean(dy*yhat)) for EVERY (m,k) row,
//         scatter-added into dz[idx], accumulated into dWxyz
#include "ptlayer.cuh"

#define TD_THREADS 256
#define TD_WARPS (TD_THREADS / 32)

static int td_grid(int m)
{
    int g = (m + TD_WARPS - 1) / TD_WARPS;
    return g < 1 ? 1 : (g > 148 * 4 ? 148 * 4 : g);
}

__global__ void k_td_rel(const float *__restrict__ p_sup, const float *__restrict__ p_qry, const int *__restrict__ idx,
                         long long rows, int k, float *__restrict__ rel)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const long long q = r / k;
        const int j = __ldg(idx + r);
        rel[3 * r] = __ldg(p_sup + 3 * j) - __ldg(p_qry + 3 * q);
        rel[3 * r + 1] = __ldg(p_sup + 3 * j + 1) - __ldg(p_qry + 3 * q + 1);
        rel[3 * r + 2] = __ldg(p_sup + 3 * j + 2) - __ldg(p_qry + 3 * q + 2);
    }
}

// MODE 0: statistics.  MODE 1: normalise + relu + max (out, argk).
template <int C, int MODE>
__global__ void __launch_bounds__(TD_THREADS) k_td_fwd(int m, int k, const float *__restrict__ rel, const int *__restrict__ idx,
                                                       const float *__restrict__ z, const float *__restrict__ wxyz,
                                                       const float *__restrict__ bn /* [4][C] */, double *__restrict__ stats,
                                                       float *__restrict__ out, unsigned char *__restrict__ argk)
{
    using M = PtMap<C>;
    constexpr int VW = M::VW, NS = M::NS;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float wa[NS][VW], wb[NS][VW], wc[NS][VW], sc[NS][VW], sh[NS][VW], s1[NS][VW], s2[NS][VW];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int ch = M::ch(lane, s, v);
            wa[s][v] = wxyz[3 * ch]; wb[s][v] = wxyz[3 * ch + 1]; wc[s][v] = wxyz[3 * ch + 2];
            sc[s][v] = MODE ? bn[ch] : 0.f; sh[s][v] = MODE ? bn[C + ch] : 0.f;
            s1[s][v] = 0.f; s2[s][v] = 0.f;
        }
    const int warps = gridDim.x * TD_WARPS;
    for (int pt = blockIdx.x * TD_WARPS + wib; pt < m; pt += warps) {
        float best[NS][VW];
        int bk[NS][VW];
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int v = 0; v < VW; v++) { best[s][v] = -3.0e38f; bk[s][v] = 0; }
        for (int kk = 0; kk < k; kk++) {
            const size_t row = (size_t)pt * k + kk;
            const int j = __ldg(idx + row);
            const float rx = __ldg(rel + 3 * row), ry = __ldg(rel + 3 * row + 1), rz = __ldg(rel + 3 * row + 2);
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float x[VW];
                pt_load<VW>(z + (size_t)j * C + M::ch(lane, s, 0), x);
#pragma unroll
                for (int v = 0; v < VW; v++) {
                    const float y = x[v] + wa[s][v] * rx + wb[s][v] * ry + wc[s][v] * rz;
                    if (MODE == 0) { s1[s][v] += y; s2[s][v] += y * y; }
                    else {
                        const float a = fmaxf(y * sc[s][v] + sh[s][v], 0.f);
                        if (a > best[s][v]) { best[s][v] = a; bk[s][v] = kk; }      // first maximum, as MaxPool1d
                    }
                }
            }
        }
        if (MODE == 1) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                pt_store<VW>(out + (size_t)pt * C + M::ch(lane, s, 0), best[s]);
#pragma unroll
                for (int v = 0; v < VW; v++) argk[(size_t)pt * C + M::ch(lane, s, v)] = (unsigned char)bk[s][v];
            }
        }
    }
    if (MODE == 0) {
        __shared__ float comb[2 * C];
        for (int i = threadIdx.x; i < 2 * C; i += TD_THREADS) comb[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int v = 0; v < VW; v++) {
                atomicAdd(&comb[M::ch(lane, s, v)], s1[s][v]);
                atomicAdd(&comb[C + M::ch(lane, s, v)], s2[s][v]);
            }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += TD_THREADS) atomicAdd(stats + i, (double)comb[i]);
    }
}

// T3: sums over the arg-max entries: sa[c] = sum_m dy, sb[c] = sum_m dy * yhat   (dy = dout where out > 0)
template <int C>
__global__ void __launch_bounds__(TD_THREADS) k_td_bwd_sums(int m, int k, const float *__restrict__ rel,
                                                            const int *__restrict__ idx, const float *__restrict__ z,
                                                            const float *__restrict__ wxyz, const float *__restrict__ bn,
                                                            const float *__restrict__ out, const unsigned char *__restrict__ argk,
                                                            const float *__restrict__ gout, double *__restrict__ sums)
{
    const int ch = blockIdx.y * TD_THREADS + threadIdx.x;
    if (ch >= C) return;
    const float wa = wxyz[3 * ch], wb = wxyz[3 * ch + 1], wc = wxyz[3 * ch + 2], mean = bn[2 * C + ch], inv = bn[3 * C + ch];
    float sa = 0.f, sb = 0.f;
    for (int pt = blockIdx.x; pt < m; pt += gridDim.x) {
        const float o = __ldg(out + (size_t)pt * C + ch);
        if (o > 0.f) {
            const float g = __ldg(gout + (size_t)pt * C + ch);
            const size_t row = (size_t)pt * k + argk[(size_t)pt * C + ch];
            const float y = __ldg(z + (size_t)__ldg(idx + row) * C + ch) + wa * __ldg(rel + 3 * row) + wb * __ldg(rel + 3 * row + 1) +
                            wc * __ldg(rel + 3 * row + 2);
            sa += g;
            sb += g * ((y - mean) * inv);
        }
    }
    atomicAdd(sums + ch, (double)sa);
    atomicAdd(sums + C + ch, (double)sb);
}

// T4: dense backward
template <int C>
__global__ void __launch_bounds__(TD_THREADS) k_td_bwd(int m, int k, const float *__restrict__ rel, const int *__restrict__ idx,
                                                       const float *__restrict__ z, const float *__restrict__ wxyz,
                                                       const float *__restrict__ bn, const float *__restrict__ coef,
                                                       const float *__restrict__ out, const unsigned char *__restrict__ argk,
                                                       const float *__restrict__ gout, float *__restrict__ gz,
                                                       float *__restrict__ gwxyz)
{
    using M = PtMap<C>;
    constexpr int VW = M::VW, NS = M::NS;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float wa[NS][VW], wb[NS][VW], wc[NS][VW], mu[NS][VW], iv[NS][VW], kk2[NS][VW], ma[NS][VW], mb[NS][VW], aw[NS][VW][3];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int ch = M::ch(lane, s, v);
            wa[s][v] = wxyz[3 * ch]; wb[s][v] = wxyz[3 * ch + 1]; wc[s][v] = wxyz[3 * ch + 2];
            mu[s][v] = bn[2 * C + ch]; iv[s][v] = bn[3 * C + ch];
            kk2[s][v] = coef[ch]; ma[s][v] = coef[C + ch]; mb[s][v] = coef[2 * C + ch];
            aw[s][v][0] = aw[s][v][1] = aw[s][v][2] = 0.f;
        }
    const int warps = gridDim.x * TD_WARPS;
    for (int pt = blockIdx.x * TD_WARPS + wib; pt < m; pt += warps) {
        float g[NS][VW];
        int ak[NS][VW];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            float o[VW];
            pt_load<VW>(gout + (size_t)pt * C + M::ch(lane, s, 0), g[s]);
            pt_load<VW>(out + (size_t)pt * C + M::ch(lane, s, 0), o);
#pragma unroll
            for (int v = 0; v < VW; v++) {
                if (!(o[v] > 0.f)) g[s][v] = 0.f;                         // relu' at the pooled entry
                ak[s][v] = argk[(size_t)pt * C + M::ch(lane, s, v)];
            }
        }
        for (int kk = 0; kk < k; kk++) {
            const size_t row = (size_t)pt * k + kk;
            const int j = __ldg(idx + row);
            const float rx = __ldg(rel + 3 * row), ry = __ldg(rel + 3 * row + 1), rz = __ldg(rel + 3 * row + 2);
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float x[VW], d[VW];
                pt_load<VW>(z + (size_t)j * C + M::ch(lane, s, 0), x);
#pragma unroll
                for (int v = 0; v < VW; v++) {
                    const float y = x[v] + wa[s][v] * rx + wb[s][v] * ry + wc[s][v] * rz;
                    const float dys = ak[s][v] == kk ? g[s][v] : 0.f;
                    d[v] = kk2[s][v] * (dys - ma[s][v] - (y - mu[s][v]) * iv[s][v] * mb[s][v]);
                    aw[s][v][0] += d[v] * rx; aw[s][v][1] += d[v] * ry; aw[s][v][2] += d[v] * rz;
                }
                pt_red_add<VW>(gz + (size_t)j * C + M::ch(lane, s, 0), d);
            }
        }
    }
    __shared__ float comb[3 * C];
    for (int i = threadIdx.x; i < 3 * C; i += TD_THREADS) comb[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int ch = M::ch(lane, s, v);
            atomicAdd(&comb[3 * ch], aw[s][v][0]); atomicAdd(&comb[3 * ch + 1], aw[s][v][1]); atomicAdd(&comb[3 * ch + 2], aw[s][v][2]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * C; i += TD_THREADS) atomicAdd(gwxyz + i, comb[i]);
}

// shared with the fused layer (ptlayer_fwd.cu / ptlayer_bwd.cu)
__global__ void k_bn_finalize(const double *__restrict__ stats, double count, int C, const float *__restrict__ gamma,
                              const float *__restrict__ beta, float *running_mean, float *running_var, float momentum, float eps,
                              int training, float *__restrict__ out);
__global__ void k_pt_bn_coef(const double *__restrict__ sums, double count, int C, const float *__restrict__ gamma,
                             const float *__restrict__ invstd, int training, float *__restrict__ coef, float *__restrict__ dgamma,
                             float *__restrict__ dbeta);

extern "C" int cb_td_rel(int m, int k, const float *p_support, const float *p_query, const int *idx, float *rel, void *stream)
{
    CB_REQUIRE(m >= 0 && k > 0 && p_support && p_query && idx && rel, CB_EINVAL, "cb_td_rel: bad arguments");
    const long long rows = (long long)m * k;
    if (rows == 0) return CB_OK;
    int g = (int)((rows + 255) / 256);
    if (g > 148 * 8) g = 148 * 8;
    k_td_rel<<<g, 256, 0, (cudaStream_t)stream>>>(p_support, p_query, idx, rows, k, rel);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_td_rel");
    return CB_OK;
}

template <int C>
static int td_forward_c(int m, int k, const float *rel, const int *idx, const float *z, const float *wxyz, const float *gamma,
                        const float *beta, float *rm, float *rv, float momentum, float eps, int training, float *out,
                        unsigned char *argk, float *bnbuf, double *stats, cudaStream_t st)
{
    cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st);
    const int grid = td_grid(m);
    if (training) k_td_fwd<C, 0><<<grid, TD_THREADS, 0, st>>>(m, k, rel, idx, z, wxyz, nullptr, stats, nullptr, nullptr);
    k_bn_finalize<<<(C + 127) / 128, 128, 0, st>>>(stats, (double)m * (double)k, C, gamma, beta, rm, rv, momentum, eps, training, bnbuf);
    k_td_fwd<C, 1><<<grid, TD_THREADS, 0, st>>>(m, k, rel, idx, z, wxyz, bnbuf, nullptr, out, argk);
    CB_COUNT(4);
    CB_CUDA_CHECK("cb_td_forward");
    return CB_OK;
}

template <int C>
static int td_backward_c(int m, int k, const float *rel, const int *idx, const float *z, const float *wxyz, const float *gamma,
                         int training, const float *bnbuf, const float *out, const unsigned char *argk, const float *gout,
                         float *gz, float *gwxyz, float *ggamma, float *gbeta, float *scratch, cudaStream_t st)
{
    double *sums = (double *)scratch;              // [2][C]
    float *coef = (float *)(sums + 2 * C);         // [3][C]
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    dim3 g3((unsigned)(m < 592 ? (m < 1 ? 1 : m) : 592), (C + TD_THREADS - 1) / TD_THREADS);
    k_td_bwd_sums<C><<<g3, TD_THREADS, 0, st>>>(m, k, rel, idx, z, wxyz, bnbuf, out, argk, gout, sums);
    k_pt_bn_coef<<<(C + 127) / 128, 128, 0, st>>>(sums, (double)m * (double)k, C, gamma, bnbuf + 3 * C, training, coef, ggamma, gbeta);
    k_td_bwd<C><<<td_grid(m), TD_THREADS, 0, st>>>(m, k, rel, idx, z, wxyz, bnbuf, coef, out, argk, gout, gz, gwxyz);
    CB_COUNT(4);
    CB_CUDA_CHECK("cb_td_backward");
    return CB_OK;
}

extern "C" int cb_td_forward(int m, int k, int c, const float *rel, const int *idx, const float *z, const float *wxyz,
                             const float *bn_weight, const float *bn_bias, float *running_mean, float *running_var,
                             float momentum, float eps, int training, float *out, unsigned char *argk, float *bnbuf,
                             double *stats, void *stream)
{
    CB_REQUIRE(m >= 0 && k >= 1 && k <= 255 && rel && idx && z && wxyz && bn_weight && bn_bias && running_mean && running_var &&
                   out && argk && bnbuf && stats, CB_EINVAL, "cb_td_forward: bad arguments");
    if (m == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (c) {
#define TD_CASE(CC) case CC: return td_forward_c<CC>(m, k, rel, idx, z, wxyz, bn_weight, bn_bias, running_mean, running_var, momentum, eps, training, out, argk, bnbuf, stats, st);
        TD_CASE(32) TD_CASE(64) TD_CASE(128) TD_CASE(256) TD_CASE(512)
#undef TD_CASE
    default:
        cb_set_error("cb_td_forward: c=%d unsupported (32,64,128,256,512)", c);
        return CB_EUNSUPPORTED;
    }
}

extern "C" int cb_td_backward(int m, int k, int c, const float *rel, const int *idx, const float *z, const float *wxyz,
                              const float *bn_weight, int training, const float *bnbuf, const float *out,
                              const unsigned char *argk, const float *grad_out, float *grad_z, float *grad_wxyz,
                              float *grad_bn_weight, float *grad_bn_bias, float *scratch, void *stream)
{
    CB_REQUIRE(m >= 0 && k >= 1 && rel && idx && z && wxyz && bn_weight && bnbuf && out && argk && grad_out && grad_z && grad_wxyz &&
                   grad_bn_weight && grad_bn_bias && scratch, CB_EINVAL, "cb_td_backward: bad arguments");
    if (m == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (c) {
#define TD_CASE(CC) case CC: return td_backward_c<CC>(m, k, rel, idx, z, wxyz, bn_weight, training, bnbuf, out, argk, grad_out, grad_z, grad_wxyz, grad_bn_weight, grad_bn_bias, scratch, st);
        TD_CASE(32) TD_CASE(64) TD_CASE(128) TD_CASE(256) TD_CASE(512)
#undef TD_CASE
    default:
        cb_set_error("cb_td_backward: c=%d unsupported", c);
        return CB_EUNSUPPORTED;
    }
}
