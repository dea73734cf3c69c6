// optim.cu — SGD with momentum and weight decay for all parameter tensors in ONE launch.
// The reference's optimiser is torch.optim.SGD(lr, momentum, weight_decay) (pytorch/tool/train.py:154); its update per element is
//     d = g + weight_decay * p;   m = momentum * m + d  (m = d on the very first step);   p = p - lr * m
// torch applies it with a multi-tensor kernel per ~32 tensors (39 launches per step for this network's 375 tensors).
// engine.GraphTrainStep already holds the step's gradient as ONE packed vector (the all-reduce buffer); parameters and
// momentum buffers stay where torch allocated them (aligned, referenced by the captured graphs) and are reached through
// a device table of pointers: element i of the packed gradient belongs to tensor t with off[t] <= i < off[t + 1].
#include "common.cuh"

__global__ void __launch_bounds__(256) k_sgd_momentum(long long total, int ntensors, const long long *__restrict__ off,
                                                     float *const *__restrict__ pp, float *const *__restrict__ mp,
                                                     const float *__restrict__ g, float lr, float momentum, float weight_decay,
                                                     int first_step)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    int t = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (!(off[t] <= i && i < off[t + 1])) {                 // binary search (rarely taken twice in a row: i grows by `stride`)
            int lo = 0, hi = ntensors - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (off[mid] <= i) lo = mid; else hi = mid - 1;
            }
            t = lo;
        }
        const long long e = i - off[t];
        float *p = pp[t], *m = mp[t];
        const float pv = p[e];
        const float d = __ldg(g + i) + weight_decay * pv;
        const float mv = first_step ? d : momentum * m[e] + d;
        m[e] = mv;
        p[e] = pv - lr * mv;
    }
}

// off: ntensors + 1 prefix offsets (elements) into g; pp / mp: ntensors device pointers each (parameter, momentum buffer)
extern "C" int cb_sgd_momentum_step(long long total, int ntensors, const long long *off, float *const *pp, float *const *mp,
                                    const float *g, float lr, float momentum, float weight_decay, int first_step, void *stream)
{
    CB_REQUIRE(total >= 0 && ntensors >= 0 && (total == 0 || (ntensors > 0 && off && pp && mp && g)), CB_EINVAL,
               "cb_sgd_momentum_step: bad arguments");
    if (total == 0) return CB_OK;
    long long blocks = (total + 256 * 4 - 1) / (256 * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    k_sgd_momentum<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(total, ntensors, off, pp, mp, g, lr, momentum, weight_decay, first_step);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_sgd_momentum_step");
    return CB_OK;
}
