// dataprep.cu — SURVEY §8(f) row 1: the step immediately BEFORE the hot path, on the device.
//   voxelize       pytorch/util/voxelize.py:38-56 (FNV64-1A voxel hash :4-16, sort, one point per voxel)
//   data_prepare   pytorch/util/data_util.py:45-92 (shift to min, voxelize, nearest-voxel_max crop around a centre,
//                  shuffle, shift to min again, feat / 255)
//   collate        pytorch/util/s3dis.py:94-130 (concatenate + cumulative int32 offsets) — clouds are written back to
//                  back into one batch buffer at a row offset that lives on the device, so a whole batch is prepared
//                  without a single device->host read.
// The reference does this in NumPy on dataloader workers (argsort over 10^5..10^6 points per cloud).
//
// Parity contract.  Bit-exact: the voxel key of every point (fp32 floor(coord / voxel_size), FNV64-1A), the set of
// occupied voxels, their counts and their output order (ascending key); the squared crop distances
// (fl(fl(dx^2 + dy^2) + dz^2), NumPy's square + sum(axis=1)); the two min-shifts; feat / 255.
// Defined modulo: WHICH point of a voxel is kept (the reference draws it from NumPy's global RNG on top of an unstable
// argsort — here: the first point in input order, or a counter-based hash of (seed, voxel)); the order among exactly
// equal crop distances (NumPy's quicksort: unspecified — here: by index); the shuffle permutation (here: sort by a
// counter-based hash).  tests/test_dataprep_gpu.py checks both halves against the reference's own functions.
#include "common.cuh"
#include <cub/cub.cuh>

typedef unsigned long long u64;

__device__ __forceinline__ u64 cb_splitmix(u64 x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// header of one cb_data_prepare call (device memory, first 256 bytes of the workspace)
struct DpHeader {
    u64 min1[3];           // order-preserving integer encoding of the per-axis minimum of the input cloud
    u64 min2[3];           // ... of the kept points (second shift)
    int nvox;              // occupied voxels
    int count;             // points kept after the crop
    int centre;            // index of the crop centre in the voxelised cloud
    int pad[5];
};

// the arithmetic runs in the dtype of `coord`, as NumPy does: float32 clouds (this repo's synthetic scenes) or float64
// (S3DIS .npy files); round-to-nearest, no contraction
template <typename T> struct DpT;
template <> struct DpT<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float big() { return 3.4e38f; }
    static __device__ __forceinline__ u64 ord(float f) { return (u64)cb_f2ord(f); }
    static __device__ __forceinline__ float unord(u64 u) { return cb_ord2f((unsigned)u); }
    static __device__ __forceinline__ u64 bits(float f) { return (u64)__float_as_uint(f); }
};
template <> struct DpT<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double big() { return 1.7e308; }
    static __device__ __forceinline__ u64 ord(double f)
    {
        const u64 u = (u64)__double_as_longlong(f);
        return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double unord(u64 u)
    {
        return __longlong_as_double((long long)((u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u));
    }
    static __device__ __forceinline__ u64 bits(double f) { return (u64)__double_as_longlong(f); }
};

__global__ void k_dp_init(DpHeader *h)
{
    if (threadIdx.x < 3) { h->min1[threadIdx.x] = ~0ull; h->min2[threadIdx.x] = ~0ull; }
    if (threadIdx.x == 3) { h->nvox = 0; h->count = 0; h->centre = 0; }
}

template <typename T> __global__ void __launch_bounds__(256) k_dp_min(const T *__restrict__ coord, int n, u64 *mn)
{
    T m[3] = {DpT<T>::big(), DpT<T>::big(), DpT<T>::big()};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int a = 0; a < 3; a++) m[a] = min(m[a], coord[3 * (size_t)i + a]);
    for (int a = 0; a < 3; a++) {
        T v = m[a];
        for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(CB_FULL_MASK, v, o));
        if ((threadIdx.x & 31) == 0) atomicMin(mn + a, DpT<T>::ord(v));
    }
}

// voxel key of every point: FNV64-1A over the three floor(coord / voxel_size) values (voxelize.py:4-16,39)
template <typename T>
__global__ void __launch_bounds__(256) k_dp_keys(const T *__restrict__ coord, int n, const u64 *mn, int shift,
                                                 T voxel_size, u64 *keys, int *vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 h = 14695981039346656037ull;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        T c = coord[3 * (size_t)i + a];
        if (shift) c = DpT<T>::sub(c, DpT<T>::unord(mn[a]));            // data_util.py:55  coord -= coord_min
        const T d = floor(DpT<T>::div(c, voxel_size));                  // voxelize.py:39   (division in coord's dtype, then floor)
        h *= 1099511628211ull;
        h ^= (u64)(long long)d;                                          // astype(uint64)
    }
    keys[i] = h;
    vals[i] = i;
}

__global__ void __launch_bounds__(256) k_dp_heads(const u64 *__restrict__ keys, int n, int *flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// voxel_start[v] = position of the first point of voxel v in the sorted order; voxel_start[nvox] = n
__global__ void __launch_bounds__(256) k_dp_starts(const int *__restrict__ flags, const int *__restrict__ scan, int n,
                                                   int *voxel_start, int *nvox_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) voxel_start[scan[i]] = i;
    if (i == n - 1) {
        const int nv = scan[i] + flags[i];
        voxel_start[nv] = n;
        *nvox_out = nv;
    }
}

// one point per voxel (voxelize.py:47-52): pick_mode 0 = first point in input order (stable sort), 1 = hashed draw
__global__ void __launch_bounds__(256) k_dp_pick(const int *__restrict__ voxel_start, const int *__restrict__ sorted_idx,
                                                 const int *nvox, int n, int pick_mode, u64 seed, int *sel, int *counts)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n || v >= *nvox) return;
    const int s = voxel_start[v], cnt = voxel_start[v + 1] - s;
    const int r = pick_mode == 0 ? 0 : (int)(cb_splitmix(seed ^ ((u64)v * 0x9E3779B97F4A7C15ull)) % (u64)cnt);
    sel[v] = sorted_idx[s + r];
    if (counts) counts[v] = cnt;
}

__global__ void k_dp_setn(DpHeader *h, int n) { h->nvox = n; }

__global__ void k_dp_centre(DpHeader *h, int centre_mode, int voxel_max, u64 seed)
{
    const int m = h->nvox;
    int c;
    if (centre_mode >= 0) c = centre_mode < m ? centre_mode : (m > 0 ? m - 1 : 0);
    else if (centre_mode == -2 && voxel_max > 0 && m > voxel_max) c = (int)(cb_splitmix(seed ^ 0xC0FFEEull) % (u64)m);   // data_util.py:59-60
    else c = m / 2;                                                                                                       // :63
    h->centre = c;
    h->count = (voxel_max > 0 && m > voxel_max) ? voxel_max : m;
}

// crop keys (data_util.py:67): d2 = sum(square(coord - coord_init), 1) in coord's dtype = fl(fl(dx*dx + dy*dy) + dz*dz);
// key = bit pattern of d2 (non-negative floats order like their bit patterns), value = j; the stable pair sort breaks
// ties by index
template <typename T>
__global__ void __launch_bounds__(256) k_dp_cropkeys(const T *__restrict__ coord, const int *__restrict__ sel,
                                                     const DpHeader *h, int n, u64 *keys, int *vals)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int m = h->nvox;
    vals[j] = j;
    if (j >= m) { keys[j] = ~0ull; return; }
    T c0[3], c[3];
    const int jc = sel[h->centre], js = sel[j];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const T mn = DpT<T>::unord(h->min1[a]);
        c0[a] = DpT<T>::sub(coord[3 * (size_t)jc + a], mn);
        c[a] = DpT<T>::sub(coord[3 * (size_t)js + a], mn);
    }
    const T dx = DpT<T>::sub(c[0], c0[0]), dy = DpT<T>::sub(c[1], c0[1]), dz = DpT<T>::sub(c[2], c0[2]);
    const T d2 = DpT<T>::add(DpT<T>::add(DpT<T>::mul(dx, dx), DpT<T>::mul(dy, dy)), DpT<T>::mul(dz, dz));
    keys[j] = DpT<T>::bits(d2);
}

// order[t] = position in the voxelised cloud of the t-th kept point, before the shuffle
__global__ void __launch_bounds__(256) k_dp_order(const int *__restrict__ cropvals, const DpHeader *h, int cropped_sorted, int n,
                                                  int shuffle, u64 seed, int *order, u64 *shufkeys)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int cnt = h->count;
    if (t >= cnt) { if (shufkeys) shufkeys[t] = ~0ull; return; }
    const int j = (cropped_sorted && h->nvox > cnt) ? cropvals[t] : t;
    order[t] = j;
    if (shufkeys) shufkeys[t] = shuffle ? ((cb_splitmix(seed ^ (0xABCDull + (u64)t * 0xD1B54A32D192ED03ull)) >> 32) << 32) | (unsigned)t : (u64)t;
}

template <typename T>
__global__ void __launch_bounds__(256) k_dp_min2(const T *__restrict__ coord, const int *__restrict__ sel,
                                                 const int *__restrict__ order, DpHeader *h, int n)
{
    T m[3] = {DpT<T>::big(), DpT<T>::big(), DpT<T>::big()};
    const int cnt = h->count;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt && t < n; t += gridDim.x * blockDim.x) {
        const int src = sel[order[t]];
        for (int a = 0; a < 3; a++) m[a] = min(m[a], DpT<T>::sub(coord[3 * (size_t)src + a], DpT<T>::unord(h->min1[a])));
    }
    for (int a = 0; a < 3; a++) {
        T v = m[a];
        for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(CB_FULL_MASK, v, o));
        if ((threadIdx.x & 31) == 0) atomicMin(&h->min2[a], DpT<T>::ord(v));
    }
}

// final gather into the batch at the device-side row offset (data_util.py:75-90, s3dis.py:117-126)
template <typename T>
__global__ void __launch_bounds__(256) k_dp_write(const T *__restrict__ coord, const T *__restrict__ feat, int fdim,
                                                  const long long *__restrict__ label, const int *__restrict__ sel,
                                                  const int *__restrict__ order, const u64 *__restrict__ shufkeys,
                                                  const DpHeader *h, int n, float feat_div, int origin_min, int *row_offset,
                                                  int out_capacity, float *out_coord, float *out_feat, long long *out_label,
                                                  int *out_index)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int cnt = h->count, base = row_offset[0];
    if (t == 0) row_offset[1] = base + cnt;
    if (t >= cnt || t >= n || base + t >= out_capacity) return;
    const int u = shufkeys ? (int)(shufkeys[t] & 0xFFFFFFFFull) : t;     // shuffled position -> kept point
    const int src = sel[order[u]];
    const size_t row = (size_t)base + t;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        T c = DpT<T>::sub(coord[3 * (size_t)src + a], DpT<T>::unord(h->min1[a]));
        if (origin_min) c = DpT<T>::sub(c, DpT<T>::unord(h->min2[a]));
        out_coord[3 * row + a] = (float)c;                                   // torch.FloatTensor(coord)   data_util.py:88
    }
    if (out_feat)
        for (int f = 0; f < fdim; f++) {
            const float v = (float)feat[(size_t)src * fdim + f];             // torch.FloatTensor(feat) / 255.   :89
            out_feat[row * fdim + f] = feat_div != 0.f ? __fdiv_rn(v, feat_div) : v;
        }
    if (out_label) out_label[row] = label ? label[src] : 0;
    if (out_index) out_index[row] = src;
}

// ---------------------------------------------------------------------------------------------
struct DpLayout {
    DpHeader *hdr;
    u64 *keys, *keys2;
    int *vals, *vals2, *flags, *scan, *voxel_start, *sel, *order;
    void *cub_tmp;
    size_t cub_bytes;
};

static size_t dp_layout(int n, char *base, DpLayout *L)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = cb_align_up(off + bytes, 256); return base ? (void *)(base + o) : (void *)nullptr; };
    const size_t N = (size_t)(n > 0 ? n : 1);
    L->hdr = (DpHeader *)take(sizeof(DpHeader));
    L->keys = (u64 *)take(8 * N);
    L->keys2 = (u64 *)take(8 * N);
    L->vals = (int *)take(4 * N);
    L->vals2 = (int *)take(4 * N);
    L->flags = (int *)take(4 * N);
    L->scan = (int *)take(4 * N);
    L->voxel_start = (int *)take(4 * (N + 1));
    L->sel = (int *)take(4 * N);
    L->order = (int *)take(4 * N);
    size_t s1 = 0, s2 = 0, s3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, s1, (u64 *)nullptr, (u64 *)nullptr, (int *)nullptr, (int *)nullptr, (int)N);
    cub::DeviceRadixSort::SortKeys(nullptr, s2, (u64 *)nullptr, (u64 *)nullptr, (int)N);
    cub::DeviceScan::ExclusiveSum(nullptr, s3, (int *)nullptr, (int *)nullptr, (int)N);
    L->cub_bytes = s1 > s2 ? (s1 > s3 ? s1 : s3) : (s2 > s3 ? s2 : s3);
    L->cub_tmp = take(L->cub_bytes + 256);
    return off;
}

extern "C" size_t cb_data_prepare_workspace_bytes(int n)
{
    DpLayout L;
    return dp_layout(n, nullptr, &L) + 256;
}

template <typename T>
static int dp_voxelize(const T *coord, int n, double voxel_size, int shift_min, const DpLayout &L, cudaStream_t st)
{
    const int g = (n + 255) / 256;
    k_dp_init<<<1, 32, 0, st>>>(L.hdr);
    if (shift_min) k_dp_min<T><<<g > 592 ? 592 : g, 256, 0, st>>>(coord, n, L.hdr->min1);
    k_dp_keys<T><<<g, 256, 0, st>>>(coord, n, L.hdr->min1, shift_min, (T)voxel_size, L.keys, L.vals);
    size_t tb = L.cub_bytes + 256;
    cub::DeviceRadixSort::SortPairs(L.cub_tmp, tb, L.keys, L.keys2, L.vals, L.vals2, n, 0, 64, st);     // stable
    k_dp_heads<<<g, 256, 0, st>>>(L.keys2, n, L.flags);
    tb = L.cub_bytes + 256;
    cub::DeviceScan::ExclusiveSum(L.cub_tmp, tb, L.flags, L.scan, n, st);
    k_dp_starts<<<g, 256, 0, st>>>(L.flags, L.scan, n, L.voxel_start, &L.hdr->nvox);
    CB_COUNT(shift_min ? 7 : 6);
    return CB_OK;
}

// voxelize(coord, voxel_size, hash_type='fnv', mode=1) (voxelize.py:38-56; the test-time call, tool/test.py):
// idx_sort (n) = input indices sorted by voxel key (stable), count (first *nvox entries) = points per voxel in key
// order, keys_sorted (n, optional) = the sorted keys.  No shift is applied (the caller shifts, as in the reference).
// coord_is_f64: the dtype of `coord` (and of the arithmetic); voxel_size is rounded to that dtype first.
extern "C" int cb_voxelize(const void *coord, int coord_is_f64, int n, double voxel_size, int *idx_sort, int *count, int *nvox,
                           unsigned long long *keys_sorted, void *workspace, size_t workspace_bytes, void *stream)
{
    CB_REQUIRE(n >= 0 && voxel_size > 0. && idx_sort && count && nvox && workspace && (coord || n == 0), CB_EINVAL,
               "cb_voxelize: bad arguments");
    CB_REQUIRE(((uintptr_t)workspace & 255) == 0, CB_EINVAL, "cb_voxelize: workspace not 256-byte aligned");
    DpLayout L;
    const size_t need = dp_layout(n, (char *)workspace, &L);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_voxelize: workspace %zu < %zu", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { cudaMemsetAsync(nvox, 0, sizeof(int), st); return CB_OK; }
    int rc = coord_is_f64 ? dp_voxelize<double>((const double *)coord, n, voxel_size, 0, L, st)
                          : dp_voxelize<float>((const float *)coord, n, voxel_size, 0, L, st);
    if (rc) return rc;
    const int g = (n + 255) / 256;
    k_dp_pick<<<g, 256, 0, st>>>(L.voxel_start, L.vals2, &L.hdr->nvox, n, 0, 0ull, L.sel, count);
    cudaMemcpyAsync(idx_sort, L.vals2, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st);
    if (keys_sorted) cudaMemcpyAsync(keys_sorted, L.keys2, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(nvox, &L.hdr->nvox, sizeof(int), cudaMemcpyDeviceToDevice, st);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_voxelize");
    return CB_OK;
}

template <typename T>
static int dp_prepare(const T *coord, const T *feat, int fdim, const long long *label, int n, double voxel_size, int voxel_max,
                      int pick_mode, int centre_mode, int shuffle, u64 seed, float feat_div, int *row_offset, int out_capacity,
                      float *out_coord, float *out_feat, long long *out_label, int *out_index, const DpLayout &L, cudaStream_t st)
{
    const int g = (n + 255) / 256;
    if (voxel_size > 0.) {
        int rc = dp_voxelize<T>(coord, n, voxel_size, 1, L, st);
        if (rc) return rc;
        k_dp_pick<<<g, 256, 0, st>>>(L.voxel_start, L.vals2, &L.hdr->nvox, n, pick_mode, seed, L.sel, nullptr);
    } else {
        // no voxelisation: every point kept, in input order.  (The reference shifts to the minimum only inside its
        // `if voxel_size:` branch; the final shift makes the result the same.)
        k_dp_init<<<1, 32, 0, st>>>(L.hdr);
        k_dp_min<T><<<g > 592 ? 592 : g, 256, 0, st>>>(coord, n, L.hdr->min1);
        k_dp_keys<T><<<g, 256, 0, st>>>(coord, n, L.hdr->min1, 1, (T)1, L.keys, L.sel);      // sel = identity (keys unused)
        k_dp_setn<<<1, 1, 0, st>>>(L.hdr, n);
    }
    k_dp_centre<<<1, 1, 0, st>>>(L.hdr, centre_mode, voxel_max, seed);
    int cropped = 0;
    if (voxel_max > 0 && n > voxel_max) {           // the crop can only trigger when the input has more points than voxel_max
        k_dp_cropkeys<T><<<g, 256, 0, st>>>(coord, L.sel, L.hdr, n, L.keys, L.vals);
        size_t tb = L.cub_bytes + 256;
        cub::DeviceRadixSort::SortPairs(L.cub_tmp, tb, L.keys, L.keys2, L.vals, L.vals2, n, 0, 64, st);   // stable: ties by index
        cropped = 1;
    }
    k_dp_order<<<g, 256, 0, st>>>(L.vals2, L.hdr, cropped, n, shuffle, seed, L.order, shuffle ? L.keys : nullptr);
    u64 *shuf = nullptr;
    if (shuffle) {
        size_t tb = L.cub_bytes + 256;
        cub::DeviceRadixSort::SortKeys(L.cub_tmp, tb, L.keys, L.keys2, n, 0, 64, st);
        shuf = L.keys2;
    }
    k_dp_min2<T><<<g > 592 ? 592 : g, 256, 0, st>>>(coord, L.sel, L.order, L.hdr, n);
    k_dp_write<T><<<g, 256, 0, st>>>(coord, feat, fdim, label, L.sel, L.order, shuf, L.hdr, n, feat_div, 1, row_offset, out_capacity,
                                     out_coord, out_feat, out_label, out_index);
    CB_COUNT(6);
    return CB_OK;
}

// data_prepare for ONE cloud, written into a batch buffer (see the file header).
//   coord (n,3), feat (n,fdim) in float32 or float64 (coord_is_f64; the arithmetic runs in that dtype), label (n) int64
//   voxel_size <= 0: no voxelisation (every point kept);  voxel_max <= 0: no crop
//   pick_mode   0 first point of the voxel | 1 hashed draw (seed)
//   centre_mode >= 0 that index of the voxelised cloud | -1 the middle point (test split) | -2 hashed draw when the cloud is
//               larger than voxel_max, else the middle (train split, data_util.py:59-63)
//   shuffle     0/1 (train.py's shuffle_index);  feat_div: 255 as in the reference, 0 = no division
//   row_offset  DEVICE int[2]: rows are written at row_offset[0]; row_offset[1] = row_offset[0] + kept count on return
//               (pass &offsets[i] of a (B+1)-long array whose first entry is 0: offsets[1:] becomes the collate offset)
//   out_*       float32 / int64 batch buffers of out_capacity rows; out_index optional: input index of every output row
extern "C" int cb_data_prepare(const void *coord, const void *feat, int coord_is_f64, int fdim, const long long *label, int n,
                               double voxel_size, int voxel_max, int pick_mode, int centre_mode, int shuffle,
                               unsigned long long seed, float feat_div, int *row_offset, int out_capacity, float *out_coord,
                               float *out_feat, long long *out_label, int *out_index, void *workspace, size_t workspace_bytes,
                               void *stream)
{
    CB_REQUIRE(n > 0 && coord && row_offset && out_coord && workspace && out_capacity > 0 && fdim >= 0, CB_EINVAL,
               "cb_data_prepare: bad arguments");
    CB_REQUIRE(((uintptr_t)workspace & 255) == 0, CB_EINVAL, "cb_data_prepare: workspace not 256-byte aligned");
    DpLayout L;
    const size_t need = dp_layout(n, (char *)workspace, &L);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_data_prepare: workspace %zu < %zu", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = coord_is_f64
                 ? dp_prepare<double>((const double *)coord, (const double *)feat, fdim, label, n, voxel_size, voxel_max, pick_mode,
                                      centre_mode, shuffle, seed, feat_div, row_offset, out_capacity, out_coord, out_feat, out_label,
                                      out_index, L, st)
                 : dp_prepare<float>((const float *)coord, (const float *)feat, fdim, label, n, voxel_size, voxel_max, pick_mode,
                                     centre_mode, shuffle, seed, feat_div, row_offset, out_capacity, out_coord, out_feat, out_label,
                                     out_index, L, st);
    if (rc) return rc;
    CB_CUDA_CHECK("cb_data_prepare");
    return CB_OK;
}
