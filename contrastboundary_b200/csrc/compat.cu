// compat.cu — the reference's OWN native ABI, served by libcbops.
//
// The reference's pybind glue (pytorch/lib/pointops/src/*/..._cuda.cpp + pointops_api.cpp) forwards
// tensor data pointers to ten `extern "C"` launchers declared in src/*/..._cuda_kernel.h.  This file
// exports those ten symbols with the reference's exact names and argument lists, so that the reference's
// unmodified glue links against libcbops.so instead of its own .cu objects (INTEGRATION.md §2;
// oracle/build_ref.sh builds that combination as oracle/_ref/dropin/pointops_cuda.so and
// tests/test_dropin_gpu.py runs the reference's model code on top of it).
//
// The reference launchers carry less information than the cb_* entry points: no stream (the reference
// launches on the legacy default stream), no support count / scene count (its kernels walk the offset
// arrays on the device) and no workspace.  The shim therefore
//   * launches on the stream set by cb_compat_set_stream (default: legacy default stream, as the reference);
//   * reads the scene count b and support count n back from the device offsets (one tiny probe kernel and
//     a stream synchronise per KNN call — the reference's Python side already synchronises per call through
//     `.item()`, pointops.py:17-20);
//   * owns one grow-only device workspace per host thread (the cb_* API stays allocation-free).
// Errors cannot be returned through a `void` launcher: they are printed to stderr and left in
// cb_last_error_string(), the outputs are then undefined (the reference does not check errors at all).
#include "common.cuh"

static thread_local void *t_ws = nullptr;
static thread_local size_t t_ws_bytes = 0;
static thread_local int *t_probe = nullptr;          // pinned, mapped host memory: {b, n}
static void *g_compat_stream = nullptr;

extern "C" void cb_compat_set_stream(void *stream) { g_compat_stream = stream; }

static void compat_fail(const char *what, int rc)
{
    fprintf(stderr, "libcbops (reference-ABI shim) %s failed (%d): %s\n", what, rc, cb_last_error_string());
}

// b = first i with new_offset[i] >= m (the cumulative ends finish at m), n = offset[b - 1]
// (the walk of get_bt_idx, knnquery_cuda_kernel.cu:51-62, done once instead of per thread)
__global__ void k_compat_probe(int m, const int *__restrict__ offset, const int *__restrict__ new_offset, int *out)
{
    int i = 0;
    while (i < (1 << 24) && new_offset[i] < m) i++;
    out[0] = i + 1;
    out[1] = offset[i];
}

static void *compat_workspace(size_t bytes, cudaStream_t st)
{
    if (bytes <= t_ws_bytes) return t_ws;
    if (t_ws) {
        cudaStreamSynchronize(st);                   // earlier launches may still read the old block
        cudaFree(t_ws);
    }
    t_ws_bytes = bytes + bytes / 4 + 4096;
    if (cudaMalloc(&t_ws, t_ws_bytes) != cudaSuccess) {
        t_ws = nullptr;
        t_ws_bytes = 0;
        (void)cudaGetLastError();
    }
    return t_ws;
}

extern "C" {

// knnquery_cuda_kernel.h:10-17.  dist2 receives SQUARED distances (pointops.py:43 takes the root).
void knnquery_cuda_launcher(int m, int nsample, const float *xyz, const float *new_xyz, const int *offset,
                            const int *new_offset, int *idx, float *dist2)
{
    if (m <= 0) return;
    cudaStream_t st = (cudaStream_t)g_compat_stream;
    if (!t_probe && cudaHostAlloc((void **)&t_probe, 2 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
        cb_set_error("cudaHostAlloc failed");
        compat_fail("knnquery_cuda_launcher", CB_ECUDA);
        return;
    }
    int *dprobe = (int *)compat_workspace(4096, st);
    if (!dprobe) { cb_set_error("cudaMalloc failed"); compat_fail("knnquery_cuda_launcher", CB_ECUDA); return; }
    k_compat_probe<<<1, 1, 0, st>>>(m, offset, new_offset, dprobe);
    cudaMemcpyAsync(t_probe, dprobe, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    const int b = t_probe[0], n = t_probe[1];
    const size_t need = cb_knn_workspace_bytes(n, m, b);
    void *ws = compat_workspace(need + 256, st);
    if (!ws) { cb_set_error("cudaMalloc of %zu bytes failed", need); compat_fail("knnquery_cuda_launcher", CB_ECUDA); return; }
    int rc = cb_knn_query(m, nsample, xyz, n, new_xyz, offset, new_offset, b, idx, dist2, 0, ws, t_ws_bytes, st);
    if (rc) compat_fail("knnquery_cuda_launcher", rc);
}

// sampling_cuda_kernel.h: n = the longest scene (pointops.py:17-20)
void furthestsampling_cuda_launcher(int b, int n, const float *xyz, const int *offset, const int *new_offset, float *tmp,
                                    int *idx)
{
    int rc = cb_furthest_sampling(b, n, xyz, offset, new_offset, tmp, idx, g_compat_stream);
    if (rc) compat_fail("furthestsampling_cuda_launcher", rc);
}

void grouping_forward_cuda_launcher(int m, int nsample, int c, const float *input, const int *idx, float *output)
{
    int rc = cb_grouping_forward(m, nsample, c, input, idx, output, g_compat_stream);
    if (rc) compat_fail("grouping_forward_cuda_launcher", rc);
}

void grouping_backward_cuda_launcher(int m, int nsample, int c, const float *grad_output, const int *idx, float *grad_input)
{
    int rc = cb_grouping_backward(m, nsample, c, grad_output, idx, grad_input, g_compat_stream);
    if (rc) compat_fail("grouping_backward_cuda_launcher", rc);
}

void interpolation_forward_cuda_launcher(int n, int c, int k, const float *input, const int *idx, const float *weight,
                                         float *output)
{
    int rc = cb_interpolation_forward(n, c, k, input, idx, weight, output, g_compat_stream);
    if (rc) compat_fail("interpolation_forward_cuda_launcher", rc);
}

void interpolation_backward_cuda_launcher(int n, int c, int k, const float *grad_output, const int *idx, const float *weight,
                                          float *grad_input)
{
    int rc = cb_interpolation_backward(n, c, k, grad_output, idx, weight, grad_input, g_compat_stream);
    if (rc) compat_fail("interpolation_backward_cuda_launcher", rc);
}

void subtraction_forward_cuda_launcher(int n, int nsample, int c, const float *input1, const float *input2, const int *idx,
                                       float *output)
{
    int rc = cb_subtraction_forward(n, nsample, c, input1, input2, idx, output, g_compat_stream);
    if (rc) compat_fail("subtraction_forward_cuda_launcher", rc);
}

void subtraction_backward_cuda_launcher(int n, int nsample, int c, const int *idx, const float *grad_output,
                                        float *grad_input1, float *grad_input2)
{
    int rc = cb_subtraction_backward(n, nsample, c, idx, grad_output, grad_input1, grad_input2, g_compat_stream);
    if (rc) compat_fail("subtraction_backward_cuda_launcher", rc);
}

void aggregation_forward_cuda_launcher(int n, int nsample, int c, int w_c, const float *input, const float *position,
                                       const float *weight, const int *idx, float *output)
{
    int rc = cb_aggregation_forward(n, nsample, c, w_c, input, position, weight, idx, output, g_compat_stream);
    if (rc) compat_fail("aggregation_forward_cuda_launcher", rc);
}

void aggregation_backward_cuda_launcher(int n, int nsample, int c, int w_c, const float *input, const float *position,
                                        const float *weight, const int *idx, const float *grad_output, float *grad_input,
                                        float *grad_position, float *grad_weight)
{
    int rc = cb_aggregation_backward(n, nsample, c, w_c, input, position, weight, idx, grad_output, grad_input, grad_position,
                                     grad_weight, g_compat_stream);
    if (rc) compat_fail("aggregation_backward_cuda_launcher", rc);
}

}   // extern "C"
