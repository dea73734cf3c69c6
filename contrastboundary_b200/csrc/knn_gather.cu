// knn_gather.cu — a1+a3 fused: grid KNN + neighbour-feature gather (the north-star kernel).
// Reference path: knnquery then feat[idx.long()] (pytorch/lib/pointops/functions/pointops.py:88-94).
//
// One persistent warp per query stream: ring-search the K nearest (knn.cuh), then move the K
// feature rows with the TMA engine — each lane that holds a neighbour index issues one
// cp.async.bulk global->shared for its C-wide row (rows land back to back in a per-warp staging
// slab), the warp waits on an mbarrier, and ONE cp.async.bulk shared->global writes the slab to
// grouped[q, k0:k0+R, :], which is contiguous.  No feature byte passes through registers; the LSU
// is left to the search.  The slab is reused only after cp.async.bulk.wait_group.read, which the
// warp issues after the NEXT query's search, so stores drain under the search.
//
// Algorithmic HBM bytes per call: 12 n + 4 n c + 8 m K + 4 m K c  (SURVEY.md §8(d)).
#include "knn.cuh"

#define KG_WARPS 8
static int g_kg_chunk_bytes = 8192;
static int g_kg_spin_ns = 0;          // back-off of a search warp that finds the hand-off queue full (0 = busy spin)
static int g_kg_mode = -1;           // -1 by row width (knn_gather_launch) | 0 one-warp TMA | 1,2,3 warp-specialised TMA rings (8/4, 12/8, 6/4)
                                     // | 4 direct register copy | 5 loader + storer warps | 6,7 one / two load-store-unit copy warps

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// L2 residency hints (bit 0: chunk stores evict_first, bit 1: row loads evict_last).  The output stream (4*N*K*C bytes) is
// written once and never read by this kernel; without a hint it pushes feature rows out of L2 before the queries of the
// next grid slab need them again once the feature table no longer fits beside it (N*C*4 > ~60 MB).
__constant__ int c_kg_l2_hint = 3;
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    if (c_kg_l2_hint & 2)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                     "l"(src), "r"(bytes), "r"(bar), "l"(l2_policy_evict_last())
                     : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, unsigned src, unsigned bytes)
{
    if (c_kg_l2_hint & 1)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes),
                     "l"(l2_policy_evict_first())
                     : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// rows per chunk: power of two <= 32 with R * row_bytes <= chunk_bytes (at least 1)
static int kg_rows_per_chunk(int c, int chunk_bytes)
{
    int r = 32;
    while (r > 1 && (size_t)r * c * 4 > (size_t)chunk_bytes) r >>= 1;
    return r;
}

// One warp per query (persistent over a query stream).  After the search the K rows are moved in
// chunks of R rows through a two-deep ring of shared-memory slabs: the lanes holding the chunk's
// neighbour indices each issue ONE cp.async.bulk global->shared (TMA) for their row, lane 0 then
// issues ONE cp.async.bulk shared->global for the whole chunk (grouped[q, e0:e0+R, :] is contiguous).
// Loads of chunk c+1 are issued before the store of chunk c has drained, and the next query's search
// runs under the tail of this query's stores; no feature byte passes through registers.
template <int KPL>
__global__ void __launch_bounds__(KG_WARPS * 32) k_knn_gather(int m, int K, int c, int R, int slab_bytes,
                                                              const float *__restrict__ new_xyz,
                                                              const float *__restrict__ feat,
                                                              const int *__restrict__ new_offset, int b, int self_query,
                                                              const CbScene *__restrict__ scenes,
                                                              const int *__restrict__ cells,
                                                              const float4 *__restrict__ sorted, int *__restrict__ idx,
                                                              float *__restrict__ dist2, float *__restrict__ grouped,
                                                              CbGridHeader *hdr, int *flagged)
{
    extern __shared__ __align__(128) unsigned char kg_smem[];
    __shared__ CbWarpScratch scratch[KG_WARPS];
    __shared__ __align__(8) unsigned long long bars[KG_WARPS][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned slab0 = smem_u32(kg_smem + (size_t)wib * 2 * slab_bytes);
    const unsigned bar0 = smem_u32(&bars[wib][0]);
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase[2] = {0u, 0u};
    const unsigned row_bytes = (unsigned)c * 4u;
    const int total_warps = gridDim.x * KG_WARPS;
    for (int w = blockIdx.x * KG_WARPS + wib; w < m; w += total_warps) {
        int q = w;
        float qx, qy, qz;
        if (self_query) {
            const float4 p = __ldg(sorted + w);
            q = __float_as_int(p.w); qx = p.x; qy = p.y; qz = p.z;
        } else {
            qx = __ldg(new_xyz + 3 * q); qy = __ldg(new_xyz + 3 * q + 1); qz = __ldg(new_xyz + 3 * q + 2);
        }
        const int s = cb_scene_of(q, new_offset, b);
        const CbScene sc = scenes[s];
        typename CbTopKSel<KPL>::type tk;
        tk.init(K, lane, sc.start);
        bool ok = cb_grid_search(tk, sc, qx, qy, qz, cells, sorted, &scratch[wib], lane);
        if (ok && tk.has_tie()) ok = false;
        if (!ok) {   // exact replay + re-gather happen in follow-up kernels
            if (lane == 0) flagged[atomicAdd(&hdr->flagged_count, 1)] = q;
            continue;
        }
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            const int e = j * 32 + lane;
            if (e < K) {
                idx[(size_t)q * K + e] = tk.out_i(j);
                dist2[(size_t)q * K + e] = tk.out_d(j);
            }
        }
        // issue the loads of chunk `ci` into slab (ci & 1)
        const int nchunks = (K + R - 1) / R;
        auto issue = [&](int ci) {
            const int e0 = ci * R;
            const int rows = min(R, K - e0);
            const unsigned bar = bar0 + 8u * (ci & 1), slab = slab0 + (unsigned)(ci & 1) * (unsigned)slab_bytes;
            if (lane == 0) mbar_expect_tx(bar, (unsigned)rows * row_bytes);
            __syncwarp();
            int my = 0;
#pragma unroll
            for (int jj = 0; jj < KPL; jj++)
                if (jj == (e0 >> 5)) my = tk.out_i(jj);
            const int rel = lane - (e0 & 31);
            if (rel >= 0 && rel < rows) bulk_g2s(slab + (unsigned)rel * row_bytes, feat + (size_t)my * c, row_bytes, bar);
        };
        if (lane == 0) bulk_wait_read0();     // both slabs free (stores of the previous query have read them)
        __syncwarp();
        issue(0);
        if (nchunks > 1) issue(1);
        for (int ci = 0; ci < nchunks; ci++) {
            const int sl = ci & 1;
            mbar_wait(bar0 + 8u * sl, phase[sl]);
            phase[sl] ^= 1u;
            const int e0 = ci * R;
            const int rows = min(R, K - e0);
            if (lane == 0) {
                bulk_s2g(grouped + ((size_t)q * K + e0) * c, slab0 + (unsigned)sl * (unsigned)slab_bytes, (unsigned)rows * row_bytes);
                if (ci + 2 < nchunks) bulk_wait_read0();   // slab `sl` must be drained before it is refilled
            }
            __syncwarp();
            if (ci + 2 < nchunks) issue(ci + 2);
        }
    }
    if (lane == 0) bulk_wait_all();
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised variant (K <= 64): 7 SEARCH warps per CTA run the ring search back to back and push
// (query, K neighbour indices) records into a shared-memory queue; 1 COPY warp pops them and drives the TMA
// engine through a ring of KGW_NS slabs with KGW_LA chunks of loads always in flight ahead of the stores.
// Search warps never wait for memory traffic, the copy warp never searches: the kernel runs at
// max(search issue time, HBM write time) instead of their per-warp sum.
// ---------------------------------------------------------------------------------------------
#define KGW_SEARCH 7
#define KGW_QN 16         // queue slots

template <int KPL, int KGW_NS, int KGW_LA>
__global__ void __launch_bounds__(256, 4) k_knn_gather_ws(int m, int K, int c, int R, int slab_bytes,
                                                       const float *__restrict__ new_xyz, const float *__restrict__ feat,
                                                       const int *__restrict__ new_offset, int b, int self_query,
                                                       const CbScene *__restrict__ scenes, const int *__restrict__ cells,
                                                       const float4 *__restrict__ sorted, int *__restrict__ idx,
                                                       float *__restrict__ dist2, float *__restrict__ grouped,
                                                       CbGridHeader *hdr, int *flagged, int spin_ns)
{
    extern __shared__ __align__(128) unsigned char kg_smem[];
    __shared__ CbWarpScratch scratch[KGW_SEARCH];
    __shared__ __align__(8) unsigned long long full_bar[KGW_NS];
    __shared__ int q_entry[KGW_QN][1 + 32 * KPL];
    __shared__ volatile int q_seq[KGW_QN];
    __shared__ int q_tail;
    __shared__ volatile int q_head;
    __shared__ volatile int producers_done;
    __shared__ int ch_q[KGW_NS], ch_e0[KGW_NS], ch_rows[KGW_NS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x < KGW_QN) q_seq[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        q_tail = 0; q_head = 0; producers_done = 0;
        for (int s = 0; s < KGW_NS; s++) mbar_init(smem_u32(&full_bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned row_bytes = (unsigned)c * 4u;
    if (wib < KGW_SEARCH) {
        // ------------------------------ search warps ------------------------------
        const int stride = gridDim.x * KGW_SEARCH;
        for (int w = blockIdx.x * KGW_SEARCH + wib; w < m; w += stride) {
            int q = w;
            float qx, qy, qz;
            if (self_query) {
                const float4 p = __ldg(sorted + w);
                q = __float_as_int(p.w); qx = p.x; qy = p.y; qz = p.z;
            } else {
                qx = __ldg(new_xyz + 3 * q); qy = __ldg(new_xyz + 3 * q + 1); qz = __ldg(new_xyz + 3 * q + 2);
            }
            const int s = cb_scene_of(q, new_offset, b);
            const CbScene sc = scenes[s];
            typename CbTopKSel<KPL>::type tk;
            tk.init(K, lane, sc.start);
            bool ok = cb_grid_search(tk, sc, qx, qy, qz, cells, sorted, &scratch[wib], lane);
            if (ok && tk.has_tie()) ok = false;
            if (!ok) {
                if (lane == 0) flagged[atomicAdd(&hdr->flagged_count, 1)] = q;
                continue;
            }
#pragma unroll
            for (int j = 0; j < KPL; j++) {
                const int e = j * 32 + lane;
                if (e < K) {
                    idx[(size_t)q * K + e] = tk.out_i(j);
                    dist2[(size_t)q * K + e] = tk.out_d(j);
                }
            }
            // enqueue (q, idx[0..K))
            int t = 0;
            if (lane == 0) {
                t = atomicAdd(&q_tail, 1);
                while (t - q_head >= KGW_QN) { if (spin_ns) __nanosleep(spin_ns); }   // queue full: wait for the copy warp
            }
            t = __shfl_sync(CB_FULL_MASK, t, 0);
            const int slot = t % KGW_QN;
#pragma unroll
            for (int j = 0; j < KPL; j++) {
                const int e = j * 32 + lane;
                if (e < K) q_entry[slot][1 + e] = tk.out_i(j);
            }
            if (lane == 0) q_entry[slot][0] = q;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) q_seq[slot] = t + 1;
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd((int *)&producers_done, 1); }
    } else {
        // ------------------------------ copy warp ------------------------------
        const unsigned slab0 = smem_u32(kg_smem);
        const int nchunks = (K + R - 1) / R;
        unsigned phase_bits = 0;          // bit s = parity to wait for on slab s
        int itemA = 0, chunkA = 0;        // issue cursor: item (queue sequence number) and chunk within it
        int gA = 0, gB = 0;               // global chunk counters: issued / stored
        bool haveA = false;               // item at itemA is available (published)
        for (;;) {
            // ---- issue loads while fewer than KGW_LA chunks are in flight
            bool progressed = false;
            while (gA - gB < KGW_LA) {
                if (!haveA) {
                    const int slot = itemA % KGW_QN;
                    if (q_seq[slot] != itemA + 1) break;                 // nothing published yet
                    __threadfence_block();
                    haveA = true;
                }
                const int slot = itemA % KGW_QN;
                const int sl = gA % KGW_NS;
                if (gA >= KGW_NS) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(KGW_NS - KGW_LA - 1) : "memory");
                    __syncwarp();
                }
                const int e0 = chunkA * R;
                const int rows = min(R, K - e0);
                const unsigned bar = smem_u32(&full_bar[sl]);
                if (lane == 0) {
                    mbar_expect_tx(bar, (unsigned)rows * row_bytes);
                    ch_q[sl] = q_entry[slot][0]; ch_e0[sl] = e0; ch_rows[sl] = rows;
                }
                __syncwarp();
                if (lane < rows)
                    bulk_g2s(slab0 + (unsigned)sl * (unsigned)slab_bytes + (unsigned)lane * row_bytes,
                             feat + (size_t)q_entry[slot][1 + e0 + lane] * c, row_bytes, bar);
                gA++;
                progressed = true;
                if (++chunkA == nchunks) {           // all chunks of this item issued: release its queue slot
                    chunkA = 0; itemA++; haveA = false;
                    __syncwarp();
                    if (lane == 0) q_head = itemA;
                }
            }
            // ---- store the oldest chunk in flight
            if (gB < gA) {
                const int sl = gB % KGW_NS;
                mbar_wait(smem_u32(&full_bar[sl]), (phase_bits >> sl) & 1u);
                phase_bits ^= 1u << sl;
                if (lane == 0)
                    bulk_s2g(grouped + ((size_t)ch_q[sl] * K + ch_e0[sl]) * c, slab0 + (unsigned)sl * (unsigned)slab_bytes,
                             (unsigned)ch_rows[sl] * row_bytes);
                __syncwarp();
                gB++;
                progressed = true;
            }
            if (!progressed) {
                // nothing in flight and nothing published: finished when every producer is done and the queue is drained
                if (producers_done == KGW_SEARCH) {
                    __threadfence_block();
                    const int tail = *((volatile int *)&q_tail);
                    if (itemA >= tail && gB == gA) break;
                }
            }
        }
        if (lane == 0) bulk_wait_all();
    }
}

// ---------------------------------------------------------------------------------------------
// LSU-copy variant (modes 6 / 7).  The TMA variants above move every 1 KB feature row with its own bulk copy: 655 360 row
// loads + 81 920 chunk stores per call = one bulk operation per ~65 cycles and SM — the rate at which the kernel saturates
// whatever the ring depth (Little's law on the ring gives the same 4.7 TB/s).  Here the copy warps use the load/store
// units instead: a queue item is one contiguous (K x c) destination block; the warp walks it as float4 vectors, KGL_U
// independent ld.global.nc (L2 hits: the feature table is resident) in flight per lane, then KGL_U streaming stores
// (st.global.cs: the 671 MB output must not evict the 42 MB table from L2).  No staging shared memory at all.
// Queue: ticket t -> slot t % KGW_QN; the slot is free for generation t / KGW_QN once q_done[slot] says so; copy warp cw
// takes the tickets t = cw (mod NCOPY).
// ---------------------------------------------------------------------------------------------
#define KGL_U 8

template <int KPL, int NCOPY>
__global__ void __launch_bounds__(256, 4) k_knn_gather_lsu(int m, int K, int c, const float *__restrict__ new_xyz,
                                                           const float *__restrict__ feat, const int *__restrict__ new_offset, int b,
                                                           int self_query, const CbScene *__restrict__ scenes,
                                                           const int *__restrict__ cells, const float4 *__restrict__ sorted,
                                                           int *__restrict__ idx, float *__restrict__ dist2,
                                                           float *__restrict__ grouped, CbGridHeader *hdr, int *flagged)
{
    constexpr int NSEARCH = 8 - NCOPY;
    __shared__ CbWarpScratch scratch[NSEARCH];
    __shared__ int q_entry[KGW_QN][1 + 32 * KPL];
    __shared__ volatile int q_seq[KGW_QN];
    __shared__ volatile int q_done[KGW_QN];
    __shared__ int q_tail;
    __shared__ volatile int producers_done;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x < KGW_QN) { q_seq[threadIdx.x] = 0; q_done[threadIdx.x] = 0; }
    if (threadIdx.x == 0) { q_tail = 0; producers_done = 0; }
    __syncthreads();
    if (wib < NSEARCH) {
        // ------------------------------ search warps ------------------------------
        const int stride = gridDim.x * NSEARCH;
        for (int w = blockIdx.x * NSEARCH + wib; w < m; w += stride) {
            int q = w;
            float qx, qy, qz;
            if (self_query) {
                const float4 p = __ldg(sorted + w);
                q = __float_as_int(p.w); qx = p.x; qy = p.y; qz = p.z;
            } else {
                qx = __ldg(new_xyz + 3 * q); qy = __ldg(new_xyz + 3 * q + 1); qz = __ldg(new_xyz + 3 * q + 2);
            }
            const int s = cb_scene_of(q, new_offset, b);
            const CbScene sc = scenes[s];
            typename CbTopKSel<KPL>::type tk;
            tk.init(K, lane, sc.start);
            bool ok = cb_grid_search(tk, sc, qx, qy, qz, cells, sorted, &scratch[wib], lane);
            if (ok && tk.has_tie()) ok = false;
            if (!ok) {
                if (lane == 0) flagged[atomicAdd(&hdr->flagged_count, 1)] = q;
                continue;
            }
#pragma unroll
            for (int j = 0; j < KPL; j++) {
                const int e = j * 32 + lane;
                if (e < K) {
                    idx[(size_t)q * K + e] = tk.out_i(j);
                    dist2[(size_t)q * K + e] = tk.out_d(j);
                }
            }
            int t = 0;
            if (lane == 0) {
                t = atomicAdd(&q_tail, 1);
                while (q_done[t % KGW_QN] != t / KGW_QN) { }              // slot still holds ticket t - KGW_QN
            }
            t = __shfl_sync(CB_FULL_MASK, t, 0);
            const int slot = t % KGW_QN;
#pragma unroll
            for (int j = 0; j < KPL; j++) {
                const int e = j * 32 + lane;
                if (e < K) q_entry[slot][1 + e] = tk.out_i(j);
            }
            if (lane == 0) q_entry[slot][0] = q;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) q_seq[slot] = t + 1;
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd((int *)&producers_done, 1); }
    } else {
        // ------------------------------ copy warps ------------------------------
        const int cw = wib - NSEARCH;
        const int c4 = c >> 2;
        const int V = K * c4;                                              // float4 vectors per item
        const float4 *feat4 = reinterpret_cast<const float4 *>(feat);
        for (int t = cw;; t += NCOPY) {
            const int slot = t % KGW_QN;
            bool finished = false;
            while (q_seq[slot] != t + 1) {
                if (producers_done == NSEARCH) {
                    __threadfence_block();
                    if (t >= *((volatile int *)&q_tail) ) { finished = true; break; }
                }
            }
            if (finished) break;
            __threadfence_block();
            const int q = q_entry[slot][0];
            float4 *dst = reinterpret_cast<float4 *>(grouped + (size_t)q * K * c);
            for (int v0 = 0; v0 < V; v0 += 32 * KGL_U) {
                float4 r[KGL_U];
#pragma unroll
                for (int u = 0; u < KGL_U; u++) {
                    const int v = v0 + u * 32 + lane;
                    if (v < V) {
                        const int row = v / c4, col = v - row * c4;
                        const int j = q_entry[slot][1 + row];
                        r[u] = __ldg(feat4 + (size_t)j * c4 + col);
                    }
                }
#pragma unroll
                for (int u = 0; u < KGL_U; u++) {
                    const int v = v0 + u * 32 + lane;
                    if (v < V) __stcs(dst + v, r[u]);
                }
            }
            __syncwarp();
            if (lane == 0) q_done[slot] = t / KGW_QN + 1;                   // the slot may take ticket t + KGW_QN
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Two-copy-warp variant (mode 5).  ncu on the variant above: 37 % of the executed instructions are the search warps
// spinning on a FULL queue — the single copy warp, which alternates blocking waits for loads (mbarrier) and for
// stores (bulk wait_group.read), is the bottleneck, not the search.  Here the copy role is split into a LOADER warp
// (pops queue items, waits for a free slab on its `empty` mbarrier, issues the row loads) and a STORER warp (waits for
// a slab's `full` mbarrier, issues the bulk store, releases the slab of the previous store once its shared-memory
// read has finished), so loads are never held up behind a draining store; 6 warps search.
// ---------------------------------------------------------------------------------------------
#define KGW2_SEARCH 6

__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int KPL, int NS>
__global__ void __launch_bounds__(256, 4) k_knn_gather_ws2(int m, int K, int c, int R, int slab_bytes,
                                                        const float *__restrict__ new_xyz, const float *__restrict__ feat,
                                                        const int *__restrict__ new_offset, int b, int self_query,
                                                        const CbScene *__restrict__ scenes, const int *__restrict__ cells,
                                                        const float4 *__restrict__ sorted, int *__restrict__ idx,
                                                        float *__restrict__ dist2, float *__restrict__ grouped,
                                                        CbGridHeader *hdr, int *flagged)
{
    extern __shared__ __align__(128) unsigned char kg_smem[];
    __shared__ CbWarpScratch scratch[KGW2_SEARCH];
    __shared__ __align__(8) unsigned long long full_bar[NS], empty_bar[NS];
    __shared__ int q_entry[KGW_QN][1 + 32 * KPL];
    __shared__ volatile int q_seq[KGW_QN];
    __shared__ int q_tail;
    __shared__ volatile int q_head;
    __shared__ volatile int producers_done;
    __shared__ volatile int total_chunks;          // -1 until the loader has issued its last chunk
    __shared__ int ch_q[NS], ch_e0[NS], ch_rows[NS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x < KGW_QN) q_seq[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        q_tail = 0; q_head = 0; producers_done = 0; total_chunks = -1;
        for (int s = 0; s < NS; s++) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned row_bytes = (unsigned)c * 4u;
    const unsigned slab0 = smem_u32(kg_smem);
    const int nchunks = (K + R - 1) / R;
    if (wib < KGW2_SEARCH) {
        // ------------------------------ search warps ------------------------------
        const int stride = gridDim.x * KGW2_SEARCH;
        for (int w = blockIdx.x * KGW2_SEARCH + wib; w < m; w += stride) {
            int q = w;
            float qx, qy, qz;
            if (self_query) {
                const float4 p = __ldg(sorted + w);
                q = __float_as_int(p.w); qx = p.x; qy = p.y; qz = p.z;
            } else {
                qx = __ldg(new_xyz + 3 * q); qy = __ldg(new_xyz + 3 * q + 1); qz = __ldg(new_xyz + 3 * q + 2);
            }
            const int s = cb_scene_of(q, new_offset, b);
            const CbScene sc = scenes[s];
            typename CbTopKSel<KPL>::type tk;
            tk.init(K, lane, sc.start);
            bool ok = cb_grid_search(tk, sc, qx, qy, qz, cells, sorted, &scratch[wib], lane);
            if (ok && tk.has_tie()) ok = false;
            if (!ok) {
                if (lane == 0) flagged[atomicAdd(&hdr->flagged_count, 1)] = q;
                continue;
            }
#pragma unroll
            for (int j = 0; j < KPL; j++) {
                const int e = j * 32 + lane;
                if (e < K) {
                    idx[(size_t)q * K + e] = tk.out_i(j);
                    dist2[(size_t)q * K + e] = tk.out_d(j);
                }
            }
            int t = 0;
            if (lane == 0) {
                t = atomicAdd(&q_tail, 1);
                while (t - q_head >= KGW_QN) { }            // queue full: wait for the loader
            }
            t = __shfl_sync(CB_FULL_MASK, t, 0);
            const int slot = t % KGW_QN;
#pragma unroll
            for (int j = 0; j < KPL; j++) {
                const int e = j * 32 + lane;
                if (e < K) q_entry[slot][1 + e] = tk.out_i(j);
            }
            if (lane == 0) q_entry[slot][0] = q;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) q_seq[slot] = t + 1;
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd((int *)&producers_done, 1); }
    } else if (wib == KGW2_SEARCH) {
        // ------------------------------ loader warp ------------------------------
        int item = 0, g = 0;
        for (;;) {
            const int slot = item % KGW_QN;
            if (q_seq[slot] != item + 1) {                   // nothing published yet
                if (producers_done == KGW2_SEARCH) {
                    __threadfence_block();
                    if (item >= *((volatile int *)&q_tail) && q_seq[slot] != item + 1) break;
                }
                continue;
            }
            __threadfence_block();
            for (int ci = 0; ci < nchunks; ci++, g++) {
                const int sl = g % NS;
                if (g >= NS) mbar_wait(smem_u32(&empty_bar[sl]), (unsigned)((g / NS - 1) & 1));   // slab released by the storer
                const int e0 = ci * R;
                const int rows = min(R, K - e0);
                const unsigned bar = smem_u32(&full_bar[sl]);
                if (lane == 0) {
                    ch_q[sl] = q_entry[slot][0]; ch_e0[sl] = e0; ch_rows[sl] = rows;
                    mbar_expect_tx(bar, (unsigned)rows * row_bytes);      // release: the storer sees ch_* after its wait
                }
                __syncwarp();
                if (lane < rows)
                    bulk_g2s(slab0 + (unsigned)sl * (unsigned)slab_bytes + (unsigned)lane * row_bytes,
                             feat + (size_t)q_entry[slot][1 + e0 + lane] * c, row_bytes, bar);
            }
            __syncwarp();
            item++;
            if (lane == 0) q_head = item;                    // queue slot free again
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); total_chunks = g; }
    } else {
        // ------------------------------ storer warp ------------------------------
        int g = 0;
        for (;;) {
            const int sl = g % NS;
            const unsigned bar = smem_u32(&full_bar[sl]);
            const unsigned parity = (unsigned)((g / NS) & 1);
            bool ready = mbar_try(bar, parity);
            if (!ready) {
                const int tc = total_chunks;
                if (tc >= 0 && g >= tc) break;               // the loader is done and every chunk has been stored
                continue;
            }
            if (lane == 0) {
                bulk_s2g(grouped + ((size_t)ch_q[sl] * K + ch_e0[sl]) * c, slab0 + (unsigned)sl * (unsigned)slab_bytes,
                         (unsigned)ch_rows[sl] * row_bytes);
                bulk_wait_read1();                           // every store but the newest has finished reading its slab
                if (g >= 1) mbar_arrive(smem_u32(&empty_bar[(g - 1) % NS]));
            }
            __syncwarp();
            g++;
        }
        if (lane == 0) bulk_wait_all();
    }
}

// ---------------------------------------------------------------------------------------------
// Direct variant: every warp searches and then copies its K rows itself with 16-byte register moves
// (U rows = up to 16 independent LDG.128 per lane in flight, streaming STG.128 to the contiguous
// grouped[q] block).  No shared-memory staging, so 32 warps per SM keep both the search issue slots and
// the memory pipes busy; the TMA variants above trade that for zero register traffic.
// ---------------------------------------------------------------------------------------------
template <int KPL, int T /* float4 per lane per row = ceil(c / 128) */>
__global__ void __launch_bounds__(256, 4) k_knn_gather_direct(int m, int K, int c, const float *__restrict__ new_xyz,
                                                              const float *__restrict__ feat,
                                                              const int *__restrict__ new_offset, int b, int self_query,
                                                              const CbScene *__restrict__ scenes,
                                                              const int *__restrict__ cells,
                                                              const float4 *__restrict__ sorted, int *__restrict__ idx,
                                                              float *__restrict__ dist2, float *__restrict__ grouped,
                                                              CbGridHeader *hdr, int *flagged)
{
    __shared__ CbWarpScratch scratch[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int cv = c >> 2;                         // float4 per row
    constexpr int U = T == 1 ? 8 : (T == 2 ? 8 : 4);     // rows in flight
    const int total_warps = gridDim.x * 8;
    for (int w = blockIdx.x * 8 + wib; w < m; w += total_warps) {
        int q = w;
        float qx, qy, qz;
        if (self_query) {
            const float4 p = __ldg(sorted + w);
            q = __float_as_int(p.w); qx = p.x; qy = p.y; qz = p.z;
        } else {
            qx = __ldg(new_xyz + 3 * q); qy = __ldg(new_xyz + 3 * q + 1); qz = __ldg(new_xyz + 3 * q + 2);
        }
        const int s = cb_scene_of(q, new_offset, b);
        const CbScene sc = scenes[s];
        typename CbTopKSel<KPL>::type tk;
        tk.init(K, lane, sc.start);
        bool ok = cb_grid_search(tk, sc, qx, qy, qz, cells, sorted, &scratch[wib], lane);
        if (ok && tk.has_tie()) ok = false;
        if (!ok) {
            if (lane == 0) flagged[atomicAdd(&hdr->flagged_count, 1)] = q;
            continue;
        }
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            const int e = j * 32 + lane;
            if (e < K) {
                idx[(size_t)q * K + e] = tk.out_i(j);
                dist2[(size_t)q * K + e] = tk.out_d(j);
            }
        }
        float4 *dst = reinterpret_cast<float4 *>(grouped + (size_t)q * K * c);
        for (int e0 = 0; e0 < K; e0 += U) {
            float4 v[U][T];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int e = e0 + u;
                int nb = 0;
#pragma unroll
                for (int j = 0; j < KPL; j++) {
                    const int t = __shfl_sync(CB_FULL_MASK, tk.out_i(j), e & 31);
                    if ((e >> 5) == j) nb = t;
                }
                const float4 *src = reinterpret_cast<const float4 *>(feat + (size_t)nb * c);
#pragma unroll
                for (int t = 0; t < T; t++) {
                    const int f = lane + 32 * t;
                    if (e < K && f < cv) v[u][t] = __ldg(src + f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int e = e0 + u;
#pragma unroll
                for (int t = 0; t < T; t++) {
                    const int f = lane + 32 * t;
                    if (e < K && f < cv) __stcs(dst + (size_t)e * cv + f, v[u][t]);
                }
            }
        }
    }
}

// generic gather of the rows of flagged queries (after the exact replay rewrote their idx)
__global__ void k_regather_flagged(int K, int c, const float *__restrict__ feat, const int *__restrict__ idx,
                                   float *__restrict__ grouped, const CbGridHeader *hdr, const int *__restrict__ flagged)
{
    const int count = hdr->flagged_count;
    for (int f = blockIdx.x; f < count; f += gridDim.x) {
        const int q = flagged[f];
        for (int e = threadIdx.x; e < K * c; e += blockDim.x) {
            const int k = e / c, ci = e - k * c;
            grouped[((size_t)q * K + k) * c + ci] = __ldg(feat + (size_t)idx[(size_t)q * K + k] * c + ci);
        }
    }
}

static int knn_gather_launch(int m, int nsample, int c, const float *xyz, const float *new_xyz, const float *feat,
                             const int *offset, const int *new_offset, int b, int n, int *idx, float *dist2, float *grouped,
                             const CbGridView &v, cudaStream_t st, bool fresh_grid = false)
{
    // copy path by row width (tools/gather_mode_probe.py, B200): rows up to 256 B go through two load/store-unit copy
    // warps per CTA (a TMA bulk copy per 128-256 B row costs more than it moves), wider rows through the TMA ring
    const int g_kg_ws = g_kg_mode >= 0 ? g_kg_mode : (c <= 64 ? 7 : 3);
    if (m == 0) return CB_OK;
    if (!fresh_grid) cb_knn_reset_flagged(v, st);      // a grid built a moment ago already has flagged_count = 0
    const int self_query = (new_xyz == xyz && m == n) ? 1 : 0;
    int slab = g_kg_chunk_bytes;
    if (slab < c * 4) slab = c * 4;
    const int R = kg_rows_per_chunk(c, slab);
    slab = R * c * 4;
    const size_t smem = (size_t)KG_WARPS * 2 * slab;
    int per_sm = (int)((200 * 1024) / (smem + 2048));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    int blocks = 148 * per_sm;
    if (blocks > (m + KG_WARPS - 1) / KG_WARPS) blocks = (m + KG_WARPS - 1) / KG_WARPS;
    if (nsample <= 64 && g_kg_ws == 4 && c <= 512) {
        int blocks_d = 148 * 4;
        if (blocks_d > (m + 7) / 8) blocks_d = (m + 7) / 8;
#define KGD_LAUNCH(KPL, T)                                                                                          \
    k_knn_gather_direct<KPL, T><<<blocks_d, 256, 0, st>>>(m, nsample, c, new_xyz, feat, new_offset, b, self_query, v.scenes, \
                                                          v.cells, v.sorted, idx, dist2, grouped, v.hdr, v.flagged)
        const int T = (c / 4 + 31) / 32;
        if (nsample <= 32) { if (T <= 1) KGD_LAUNCH(1, 1); else if (T == 2) KGD_LAUNCH(1, 2); else KGD_LAUNCH(1, 4); }
        else { if (T <= 1) KGD_LAUNCH(2, 1); else if (T == 2) KGD_LAUNCH(2, 2); else KGD_LAUNCH(2, 4); }
#undef KGD_LAUNCH
        cb_knn_replay_launch(nsample, m, xyz, new_xyz, offset, new_offset, b, idx, dist2, 0, v, st);
        k_regather_flagged<<<148, 256, 0, st>>>(nsample, c, feat, idx, grouped, v.hdr, v.flagged);
        CB_COUNT(4);
        CB_CUDA_CHECK("cb_knn_gather");
        return CB_OK;
    }
    if (nsample <= 64 && (g_kg_ws == 6 || g_kg_ws == 7)) {
        int blocks_l = 148 * 4;
        const int ns_w = g_kg_ws == 6 ? 7 : 6;
        if (blocks_l > (m + ns_w - 1) / ns_w) blocks_l = (m + ns_w - 1) / ns_w;
#define KGL_LAUNCH(KPL, NC)                                                                                          \
    k_knn_gather_lsu<KPL, NC><<<blocks_l, 256, 0, st>>>(m, nsample, c, new_xyz, feat, new_offset, b, self_query, v.scenes, \
                                                        v.cells, v.sorted, idx, dist2, grouped, v.hdr, v.flagged)
        if (nsample <= 32) { if (g_kg_ws == 6) KGL_LAUNCH(1, 1); else KGL_LAUNCH(1, 2); }
        else { if (g_kg_ws == 6) KGL_LAUNCH(2, 1); else KGL_LAUNCH(2, 2); }
#undef KGL_LAUNCH
        cb_knn_replay_launch(nsample, m, xyz, new_xyz, offset, new_offset, b, idx, dist2, 0, v, st);
        k_regather_flagged<<<148, 256, 0, st>>>(nsample, c, feat, idx, grouped, v.hdr, v.flagged);
        CB_COUNT(3);
        CB_CUDA_CHECK("cb_knn_gather");
        return CB_OK;
    }
    if (nsample <= 64 && g_kg_ws == 5) {
        const int ns = 6;
        const size_t smem_ws = (size_t)ns * slab;
        int per = (int)(233472 / (smem_ws + 6800 + 1024));
        if (per > 4) per = 4;
        if (per < 1) per = 1;
        int blocks_ws = 148 * per;
        if (blocks_ws > (m + KGW2_SEARCH - 1) / KGW2_SEARCH) blocks_ws = (m + KGW2_SEARCH - 1) / KGW2_SEARCH;
        if (nsample <= 32) {
            cudaFuncSetAttribute(k_knn_gather_ws2<1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws);
            k_knn_gather_ws2<1, 6><<<blocks_ws, 256, smem_ws, st>>>(m, nsample, c, R, slab, new_xyz, feat, new_offset, b, self_query,
                                                                  v.scenes, v.cells, v.sorted, idx, dist2, grouped, v.hdr, v.flagged);
        } else {
            cudaFuncSetAttribute(k_knn_gather_ws2<2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws);
            k_knn_gather_ws2<2, 6><<<blocks_ws, 256, smem_ws, st>>>(m, nsample, c, R, slab, new_xyz, feat, new_offset, b, self_query,
                                                                  v.scenes, v.cells, v.sorted, idx, dist2, grouped, v.hdr, v.flagged);
        }
        cb_knn_replay_launch(nsample, m, xyz, new_xyz, offset, new_offset, b, idx, dist2, 0, v, st);
        k_regather_flagged<<<148, 256, 0, st>>>(nsample, c, feat, idx, grouped, v.hdr, v.flagged);
        CB_COUNT(4);
        CB_CUDA_CHECK("cb_knn_gather");
        return CB_OK;
    }
    if (nsample <= 64 && g_kg_ws) {
        // warp-specialised kernel: ring of NS slabs, LA chunks in flight (mode 1: 8/4, mode 2: 12/8, mode 3: 6/4)
        const int ns = g_kg_ws == 2 ? 12 : (g_kg_ws == 3 ? 6 : 8);
        const size_t smem_ws = (size_t)ns * slab;
        int per = (int)(233472 / (smem_ws + 6400 + 1024));     // 228 KB of shared memory per SM, 1 KB reserved per CTA
        if (per > 4) per = 4;
        if (per < 1) per = 1;
        int blocks_ws = 148 * per;
        if (blocks_ws > (m + KGW_SEARCH - 1) / KGW_SEARCH) blocks_ws = (m + KGW_SEARCH - 1) / KGW_SEARCH;
#define KGW_LAUNCH(KPL, NS, LA)                                                                                      \
    do {                                                                                                             \
        cudaFuncSetAttribute(k_knn_gather_ws<KPL, NS, LA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws); \
        k_knn_gather_ws<KPL, NS, LA><<<blocks_ws, 256, smem_ws, st>>>(m, nsample, c, R, slab, new_xyz, feat, new_offset, b, \
                                                                      self_query, v.scenes, v.cells, v.sorted, idx, dist2,  \
                                                                      grouped, v.hdr, v.flagged, g_kg_spin_ns);            \
    } while (0)
        if (nsample <= 32) {
            if (g_kg_ws == 2) KGW_LAUNCH(1, 12, 8); else if (g_kg_ws == 3) KGW_LAUNCH(1, 6, 4); else KGW_LAUNCH(1, 8, 4);
        } else {
            if (g_kg_ws == 2) KGW_LAUNCH(2, 12, 8); else if (g_kg_ws == 3) KGW_LAUNCH(2, 6, 4); else KGW_LAUNCH(2, 8, 4);
        }
#undef KGW_LAUNCH
        cb_knn_replay_launch(nsample, m, xyz, new_xyz, offset, new_offset, b, idx, dist2, 0, v, st);
        k_regather_flagged<<<148, 256, 0, st>>>(nsample, c, feat, idx, grouped, v.hdr, v.flagged);
        CB_COUNT(4);
        CB_CUDA_CHECK("cb_knn_gather");
        return CB_OK;
    }
#define KG_LAUNCH(KPL)                                                                                              \
    do {                                                                                                            \
        cudaFuncSetAttribute(k_knn_gather<KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        k_knn_gather<KPL><<<blocks, KG_WARPS * 32, smem, st>>>(m, nsample, c, R, slab, new_xyz, feat, new_offset, b, \
                                                               self_query, v.scenes, v.cells, v.sorted, idx, dist2, \
                                                               grouped, v.hdr, v.flagged);                          \
    } while (0)
    if (nsample <= 32) KG_LAUNCH(1);
    else if (nsample <= 64) KG_LAUNCH(2);
    else if (nsample <= 128) KG_LAUNCH(4);
    else KG_LAUNCH(8);
#undef KG_LAUNCH
    cb_knn_replay_launch(nsample, m, xyz, new_xyz, offset, new_offset, b, idx, dist2, 0, v, st);
    k_regather_flagged<<<148, 256, 0, st>>>(nsample, c, feat, idx, grouped, v.hdr, v.flagged);
    CB_COUNT(4);
    CB_CUDA_CHECK("cb_knn_gather");
    return CB_OK;
}

static bool kg_tma_ok(int c, int nsample, const float *feat, const float *grouped)
{
    return (c % 4 == 0) && (((uintptr_t)feat | (uintptr_t)grouped) % 16 == 0) && nsample <= 256 && (size_t)c * 4 <= 16384;
}

extern "C" int cb_knn_gather_set_spin_ns(int ns) { if (ns >= 0 && ns <= 4096) g_kg_spin_ns = ns; return g_kg_spin_ns; }

extern "C" int cb_knn_gather_set_l2_hint(int bits)
{
    bits &= 3;
    return cudaMemcpyToSymbol(c_kg_l2_hint, &bits, sizeof(int)) == cudaSuccess ? bits : -1;
}
extern "C" int cb_knn_gather_set_mode(int mode) { g_kg_mode = (mode >= -1 && mode <= 7) ? mode : -1; return g_kg_mode; }

extern "C" int cb_knn_gather_set_chunk_bytes(int bytes)
{
    if (bytes >= 512 && bytes <= 16384) g_kg_chunk_bytes = bytes;
    return g_kg_chunk_bytes;
}

// split form: the support grid was built by cb_grid_build (same workspace)
extern "C" int cb_knn_gather_grid(int m, int nsample, int c, const float *xyz, int n, const float *new_xyz, const float *feat,
                                  const int *offset, const int *new_offset, int b, int *idx, float *dist2, float *grouped,
                                  void *grid, size_t grid_bytes, void *stream)
{
    CB_REQUIRE(m >= 0 && n >= 0 && b > 0 && c > 0 && nsample >= 1 && nsample <= 256, CB_EINVAL, "cb_knn_gather_grid: bad sizes");
    CB_REQUIRE((xyz || n == 0) && feat && offset && new_offset && grid && idx && dist2 && grouped, CB_EINVAL,
               "cb_knn_gather_grid: NULL pointer");
    if (!new_xyz) new_xyz = xyz;
    CB_REQUIRE(kg_tma_ok(c, nsample, feat, grouped), CB_EUNSUPPORTED, "cb_knn_gather_grid: needs c %% 4 == 0 and 16-byte aligned rows");
    CbGridView v;
    const size_t need = cb_grid_layout(n, m, b, grid, &v);
    CB_REQUIRE(grid_bytes >= need, CB_EWORKSPACE, "cb_knn_gather_grid: workspace %zu < %zu", grid_bytes, need);
    return knn_gather_launch(m, nsample, c, xyz, new_xyz, feat, offset, new_offset, b, n, idx, dist2, grouped, v,
                             (cudaStream_t)stream);
}

extern "C" int cb_knn_gather(int m, int nsample, int c, const float *xyz, int n, const float *new_xyz, const float *feat,
                             const int *offset, const int *new_offset, int b, int *idx, float *dist2, float *grouped,
                             void *workspace, size_t workspace_bytes, void *stream)
{
    CB_REQUIRE(m >= 0 && n >= 0 && b > 0 && c > 0, CB_EINVAL, "cb_knn_gather: bad sizes");
    CB_REQUIRE(nsample >= 1 && nsample <= CB_KNN_MAX_NSAMPLE, CB_EINVAL, "cb_knn_gather: nsample=%d", nsample);
    CB_REQUIRE((xyz || n == 0) && feat && offset && new_offset && workspace && idx && dist2 && grouped, CB_EINVAL,
               "cb_knn_gather: NULL pointer");
    CB_REQUIRE(((uintptr_t)workspace & 255) == 0, CB_EINVAL, "cb_knn_gather: workspace not 256-byte aligned");
    if (!new_xyz) new_xyz = xyz;
    cudaStream_t st = (cudaStream_t)stream;
    if (!kg_tma_ok(c, nsample, feat, grouped)) {   // unfused composition with identical results
        int rc = cb_knn_query(m, nsample, xyz, n, new_xyz, offset, new_offset, b, idx, dist2, 0, workspace,
                              workspace_bytes, stream);
        if (rc) return rc;
        return cb_grouping_forward(m, nsample, c, feat, idx, grouped, stream);
    }
    CbGridView v;
    const size_t need = cb_grid_layout(n, m, b, workspace, &v);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_knn_gather: workspace %zu < %zu", workspace_bytes, need);
    int rc = cb_grid_build_impl(xyz, n, offset, b, nsample, v, st);
    if (rc) return rc;
    return knn_gather_launch(m, nsample, c, xyz, new_xyz, feat, offset, new_offset, b, n, idx, dist2, grouped, v, st, true);
}
