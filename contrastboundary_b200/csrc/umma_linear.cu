// umma_linear.cu — tall-skinny FP32 linear layers on the 5th-generation tensor cores (Blackwell, sm_100a):
//     Y (n x N) = A (n x K) . B^T (+ bias)      B = W (N x K, forward) or W^T (dgrad: W is K x N)
// tcgen05.mma kind::tf32 issued by ONE thread per CTA, operands in shared memory (K-major core-matrix layout, no
// swizzle, written by the CTA's threads while they split every FP32 value into TF32 hi + lo), accumulator in TENSOR
// MEMORY (TMEM, 128 lanes x N columns), completion through tcgen05.commit -> mbarrier, epilogue tcgen05.ld -> registers
// -> + bias -> global.  3xTF32 error compensation (hi*hi + lo*hi + hi*lo, FP32 accumulate) keeps the FP32 parity of
// tc_gemm.cu (1e-5 against float64) — the reference runs these layers as FP32 cuBLAS GEMMs (nn.Linear in
// pytorch/model/blocks.py:33,72,76,108,127-131).
// The legacy warp-level path (mma.sync, tc_gemm.cu) stays as the fallback for shapes this kernel does not take
// (K % 8 != 0, N % 16 != 0) and behind cb_linear_set_umma(0).
#include "common.cuh"

#define UM_THREADS 128
#define UM_KC 32                       // K floats staged per chunk (4 MMA k-steps of 8)

__device__ __forceinline__ unsigned um_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void um_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void um_mbar_wait(unsigned bar, unsigned parity)
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

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): in 16-byte units
//   bits [0,14) start address, [16,30) leading byte offset (between the two 16-byte K chunks of one MMA),
//   [32,46) stride byte offset (between 8-row groups), [46,48) version = 1, [61,64) layout type = 0
__device__ __forceinline__ unsigned long long um_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    return (unsigned long long)((smem_addr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// D[tmem] (+)= A[smem] . B[smem]^T, one 128 x N x 8 TF32 MMA (cute::SM100_MMA_TF32_SS)
__device__ __forceinline__ void um_mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                            unsigned accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void um_split(float x, float &hi, float &lo)
{
    unsigned h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));          // round to nearest TF32 (10-bit mantissa)
    hi = __uint_as_float(h);
    lo = x - hi;                                                   // exact in FP32; the MMA reads its TF32 part
}

// element (row, kf) of a staged tile, kf in [0, UM_KC): core matrix (row / 8, kf / 4) of 8 rows x 16 bytes
__device__ __forceinline__ unsigned um_off(int row, int kf)
{
    return (unsigned)(((row >> 3) * (UM_KC / 4) + (kf >> 2)) * 128 + (row & 7) * 16 + (kf & 3) * 4);
}

template <int TRANS_B>
__global__ void __launch_bounds__(UM_THREADS) k_umma_linear(int n, int K, int N, int ncols, const float *__restrict__ A, int lda,
                                                            const float *__restrict__ W, const float *__restrict__ bias,
                                                            float *__restrict__ Y, int ldy, int n0 /* first output column */,
                                                            int ldw)
{
    extern __shared__ __align__(1024) unsigned char um_smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned a_hi = um_smem_u32(um_smem), a_lo = a_hi + 128 * UM_KC * 4;
    const unsigned b_hi = a_lo + 128 * UM_KC * 4, b_lo = b_hi + (unsigned)N * UM_KC * 4;
    unsigned char *pa_hi = um_smem, *pa_lo = pa_hi + 128 * UM_KC * 4, *pb_hi = pa_lo + 128 * UM_KC * 4,
                  *pb_lo = pb_hi + (size_t)N * UM_KC * 4;
    const unsigned bar = um_smem_u32(&mbar);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(&tmem_slot)), "r"((unsigned)ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        um_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    const long long row0 = (long long)blockIdx.x * 128;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
    unsigned phase = 0;
    for (int k0 = 0; k0 < K; k0 += UM_KC) {
        const int kc = min(UM_KC, K - k0);                 // multiple of 8
        // ---- stage the A chunk: lane -> (row % 8, 16-byte chunk % 4): 64 contiguous bytes of 8 rows per warp instruction,
        //      128 contiguous bytes of shared memory per 8 lanes (conflict free)
        for (int it = warp; it < 16 * (UM_KC / 16); it += UM_THREADS / 32) {
            const int g = it % 16, cq = it / 16;                       // 8-row group, quad of 16-byte chunks
            const int row = g * 8 + (lane & 7), c = cq * 4 + (lane >> 3);
            const int kf = c * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kf < kc && row0 + row < n) v = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(row0 + row) * lda + k0 + kf));
            float4 h, l;
            um_split(v.x, h.x, l.x); um_split(v.y, h.y, l.y); um_split(v.z, h.z, l.z); um_split(v.w, h.w, l.w);
            const unsigned o = um_off(row, kf);
            *reinterpret_cast<float4 *>(pa_hi + o) = h;
            *reinterpret_cast<float4 *>(pa_lo + o) = l;
        }
        // ---- stage the B chunk (N rows)
        if (!TRANS_B) {                                     // W is (N x K): same mapping as A
            const int groups = N / 8;
            for (int it = warp; it < groups * (UM_KC / 16); it += UM_THREADS / 32) {
                const int g = it % groups, cq = it / groups;
                const int row = g * 8 + (lane & 7), c = cq * 4 + (lane >> 3);
                const int kf = c * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kf < kc) v = __ldg(reinterpret_cast<const float4 *>(W + (size_t)(n0 + row) * ldw + k0 + kf));
                float4 h, l;
                um_split(v.x, h.x, l.x); um_split(v.y, h.y, l.y); um_split(v.z, h.z, l.z); um_split(v.w, h.w, l.w);
                const unsigned o = um_off(row, kf);
                *reinterpret_cast<float4 *>(pb_hi + o) = h;
                *reinterpret_cast<float4 *>(pb_lo + o) = l;
            }
        } else {                                            // W is (K x N): B(row, kf) = W[k0 + kf][n0 + row]
            for (int e = tid; e < N * UM_KC; e += UM_THREADS) {
                const int row = e % N, kf = e / N;
                float v = 0.f;
                if (kf < kc) v = __ldg(W + (size_t)(k0 + kf) * ldw + n0 + row);
                float h, l;
                um_split(v, h, l);
                const unsigned o = um_off(row, kf);
                *reinterpret_cast<float *>(pb_hi + o) = h;
                *reinterpret_cast<float *>(pb_lo + o) = l;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned lbo = 128, sbo = (UM_KC / 4) * 128;
            for (int s = 0; s < kc / 8; s++) {
                const unsigned adv = (unsigned)s * 256;                   // two 16-byte K chunks per MMA
                const unsigned long long dah = um_desc(a_hi + adv, lbo, sbo), dal = um_desc(a_lo + adv, lbo, sbo);
                const unsigned long long dbh = um_desc(b_hi + adv, lbo, sbo), dbl = um_desc(b_lo + adv, lbo, sbo);
                um_mma_tf32(tmem, dah, dbh, idesc, (k0 > 0 || s > 0) ? 1u : 0u);
                um_mma_tf32(tmem, dal, dbh, idesc, 1u);
                um_mma_tf32(tmem, dah, dbl, idesc, 1u);
            }
            // completion of every MMA issued so far -> one arrival on the mbarrier (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        um_mbar_wait(bar, phase);                           // the staged chunk may be overwritten / the accumulator read
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- epilogue: warp w owns TMEM lanes 32w .. 32w+31 = output rows; 8 columns per tcgen05.ld
    const long long row = row0 + warp * 32 + lane;
    for (int col = 0; col < N; col += 8) {
        unsigned r[8];
        const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)col;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < n) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = __uint_as_float(r[e]) + (bias ? __ldg(bias + n0 + col + e) : 0.f);
            float4 *dst = reinterpret_cast<float4 *>(Y + (size_t)row * ldy + n0 + col);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)ncols));
}

// ---------------------------------------------------------------------------------------------
// v2: persistent, warp-specialised, pipelined.  One CTA per SM walks the 128-row tiles:
//   warps 0-3  epilogue   TMEM -> registers (tcgen05.ld) -> + bias -> global      (warp w <-> TMEM lanes 32w..32w+31)
//   warps 4-15 producers  global A rows -> TF32 hi / lo -> shared memory core-matrix layout; one group of 4 warps per slot of
//                         the UM2_STAGES-deep ring, so UM2_STAGES chunks (48 KB of loads per SM) are in flight at once —
//                         the kernel is an HBM stream and bytes in flight are what buys bandwidth (Little's law)
//   warp  16   one lane issues every tcgen05.mma; tcgen05.commit releases ring slots / publishes accumulators
// W (hi / lo, all of K) is staged ONCE per CTA.  Two accumulators in TMEM (2 x N columns): the MMAs of tile t+1 run while
// the epilogue drains tile t.  Four mbarrier families: full[s] (4 producer warps -> MMA), empty[s] (commit -> producers),
// tfull[b] (commit -> epilogue), tempty[b] (4 epilogue warps -> MMA).
// ---------------------------------------------------------------------------------------------
#define UM2_STAGES 3                           // v2: three producer groups, loads held in registers
#define UM3_STAGES 2                           // v3: two producer groups ...
#define UM3_DEPTH 3                            //     ... each with a 3-deep cp.async ring of raw FP32 chunks (96 KB in flight per SM)
#define UM_THREADS_OF(STAGES) (32 * (4 + 4 * (STAGES) + 1))
#define UM2_THREADS UM_THREADS_OF(UM2_STAGES)
#define UM3_THREADS UM_THREADS_OF(UM3_STAGES)
#define UM2_EPI_PITCH 36                      // floats per staged epilogue row (32 + 4: conflict-free float4 rows)
#define UM2_EPI_BYTES (4 * 32 * UM2_EPI_PITCH * 4)

__device__ __forceinline__ void um_mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// v3 = the same pipeline with DEPTH > 0: the producer threads do not hold their loads in registers (3 groups x 128 threads x 8
// float4 = 48 KB per SM, the measured limit of v2: HBM-latency bound at 35 % DRAM utilisation) but stream them with
// cp.async.cg into a private DEPTH-deep ring of raw FP32 chunks in shared memory; every thread later reads back exactly the 16-byte
// pieces it copied itself (no barrier between copy and use, only cp.async.wait_group), splits them into TF32 hi / lo and
// writes the core-matrix tiles as before.  STAGES x DEPTH x 16 KB are in flight per SM.
__device__ __forceinline__ void um_cp_async16(unsigned dst, const void *src, unsigned src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int TRANS_B, int STAGES, int DEPTH>
__global__ void __launch_bounds__(UM_THREADS_OF(STAGES), 1) k_umma_linear2(int n, int K, int nb, int ncols, int ntiles, const float *__restrict__ A,
                                                                 int lda, const float *__restrict__ W, const float *__restrict__ bias,
                                                                 float *__restrict__ Y, int ldy, int Ntot, int ldw)
{
    // blockIdx.y = column block of nb output columns (the last one may be narrower, still a multiple of 16)
    const int n0 = blockIdx.y * nb;
    const int N = min(nb, Ntot - n0);
    extern __shared__ __align__(1024) unsigned char um_smem[];
    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 4];
    constexpr int MMA_WARP = 4 + 4 * STAGES, NTHREADS = UM_THREADS_OF(STAGES);
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunk = (K + UM_KC - 1) / UM_KC;
    const unsigned w_bytes = (unsigned)N * UM_KC * 4;                 // one K chunk of W (hi or lo)
    // layout: [W hi chunks][W lo chunks][stage 0: A hi, A lo][stage 1 ...]
    unsigned char *pw_hi = um_smem, *pw_lo = pw_hi + (size_t)nchunk * w_bytes;
    unsigned char *pa = pw_lo + (size_t)nchunk * w_bytes;
    unsigned char *pstage = pa + (size_t)STAGES * 2 * 128 * UM_KC * 4;        // epilogue transposition buffers (4 warps)
    unsigned char *praw = pstage + UM2_EPI_BYTES;                                 // DEPTH > 0: raw FP32 chunk rings of the producer groups
    const unsigned sw_hi = um_smem_u32(pw_hi), sw_lo = um_smem_u32(pw_lo), sa = um_smem_u32(pa);
    const unsigned a_stage = 2 * 128 * UM_KC * 4;
    const unsigned b_full = um_smem_u32(&bars[0]), b_empty = um_smem_u32(&bars[STAGES]), b_tfull = um_smem_u32(&bars[2 * STAGES]),
                   b_tempty = um_smem_u32(&bars[2 * STAGES + 2]);
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(&tmem_slot)), "r"((unsigned)ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { um_mbar_init(b_full + 8 * s, 4); um_mbar_init(b_empty + 8 * s, 1); }
        for (int b = 0; b < 2; b++) { um_mbar_init(b_tfull + 8 * b, 1); um_mbar_init(b_tempty + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- W, once: every thread helps
    for (int j = 0; j < nchunk; j++) {
        const int k0 = j * UM_KC, kc = min(UM_KC, K - k0);
        unsigned char *dh = pw_hi + (size_t)j * w_bytes, *dl = pw_lo + (size_t)j * w_bytes;
        if (!TRANS_B) {
            const int groups = N / 8;
            for (int it = warp; it < groups * (UM_KC / 16); it += NTHREADS / 32) {
                const int g = it % groups, cq = it / groups;
                const int row = g * 8 + (lane & 7), kf = (cq * 4 + (lane >> 3)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kf < kc) v = __ldg(reinterpret_cast<const float4 *>(W + (size_t)(n0 + row) * ldw + k0 + kf));
                float4 h, l;
                um_split(v.x, h.x, l.x); um_split(v.y, h.y, l.y); um_split(v.z, h.z, l.z); um_split(v.w, h.w, l.w);
                const unsigned o = um_off(row, kf);
                *reinterpret_cast<float4 *>(dh + o) = h;
                *reinterpret_cast<float4 *>(dl + o) = l;
            }
        } else {                                            // W is (K x N): 4 consecutive output columns per thread
            const int nq = N / 4;
            for (int e = tid; e < nq * UM_KC; e += NTHREADS) {
                const int row = (e % nq) * 4, kf = e / nq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kf < kc) v = __ldg(reinterpret_cast<const float4 *>(W + (size_t)(k0 + kf) * ldw + n0 + row));
                const float va[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float h, l;
                    um_split(va[q], h, l);
                    const unsigned o = um_off(row + q, kf);
                    *reinterpret_cast<float *>(dh + o) = h;
                    *reinterpret_cast<float *>(dl + o) = l;
                }
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned lbo = 128, sbo = (UM_KC / 4) * 128;

    if (warp >= 4 && warp < MMA_WARP) {
        // ===== producers: group `stage` (4 warps) fills ring slot `stage`, i.e. the chunks c = stage, stage + STAGES, ... of
        //       this CTA's chunk sequence (tile-major) =====
        const int stage = (warp - 4) / 4, pw = (warp - 4) % 4;
        int my_tiles = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) my_tiles++;
        const long long total_chunks = (long long)my_tiles * nchunk;
        unsigned char *dh = pa + (size_t)stage * a_stage, *dl = dh + 128 * UM_KC * 4;
        unsigned use = 0;
        if (DEPTH == 0) {
            for (long long c = stage; c < total_chunks; c += STAGES, use++) {
                const int tseq = (int)(c / nchunk), j = (int)(c % nchunk);
                const long long row0 = ((long long)blockIdx.x + (long long)tseq * gridDim.x) * 128;
                const int k0 = j * UM_KC, kc = min(UM_KC, K - k0);
                // the loads do not depend on the slot: issue them before waiting for it
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int it = pw + 4 * u;
                    const int g = it % 16, cq = it / 16;
                    const int row = g * 8 + (lane & 7), kf = (cq * 4 + (lane >> 3)) * 4;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (kf < kc && row0 + row < n) v[u] = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(row0 + row) * lda + k0 + kf));
                }
                um_mbar_wait(b_empty + 8 * stage, (use & 1u) ^ 1u);              // slot free (passes on first use)
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int it = pw + 4 * u;
                    const int g = it % 16, cq = it / 16;
                    const int row = g * 8 + (lane & 7), kf = (cq * 4 + (lane >> 3)) * 4;
                    float4 h, l;
                    um_split(v[u].x, h.x, l.x); um_split(v[u].y, h.y, l.y); um_split(v[u].z, h.z, l.z); um_split(v[u].w, h.w, l.w);
                    const unsigned o = um_off(row, kf);
                    *reinterpret_cast<float4 *>(dh + o) = h;
                    *reinterpret_cast<float4 *>(dl + o) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) um_mbar_arrive(b_full + 8 * stage);
            }
        } else {
            // private ring: slab d of this group, piece u of this thread at ((d * 8 + u) * 128 + thread-in-group) * 16 bytes
            constexpr int DD = DEPTH > 0 ? DEPTH : 1;
            const int tig = pw * 32 + lane;
            unsigned char *myraw = praw + ((size_t)stage * DD * 8 * 128 + tig) * 16;
            const unsigned myraw_s = um_smem_u32(myraw);
            auto issue = [&](long long c, int slab) {
                if (c < total_chunks) {
                    const int tseq = (int)(c / nchunk), j = (int)(c % nchunk);
                    const long long row0 = ((long long)blockIdx.x + (long long)tseq * gridDim.x) * 128;
                    const int k0 = j * UM_KC, kc = min(UM_KC, K - k0);
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int it = pw + 4 * u;
                        const int g = it % 16, cq = it / 16;
                        const int row = g * 8 + (lane & 7), kf = (cq * 4 + (lane >> 3)) * 4;
                        const bool in = kf < kc && row0 + row < n;
                        const float *src = in ? A + (size_t)(row0 + row) * lda + k0 + kf : A;
                        um_cp_async16(myraw_s + (unsigned)((slab * 8 + u) * 128 * 16), src, in ? 16u : 0u);   // src-size 0: zero fill
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");           // always: keeps the group count uniform
            };
#pragma unroll
            for (int d = 0; d < DD; d++) issue((long long)stage + (long long)d * STAGES, d);
            int slab = 0;
            for (long long c = stage; c < total_chunks; c += STAGES, use++) {
                asm volatile("cp.async.wait_group %0;" ::"n"(DD - 1) : "memory");
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = *reinterpret_cast<const float4 *>(myraw + (size_t)((slab * 8 + u) * 128 * 16));
                um_mbar_wait(b_empty + 8 * stage, (use & 1u) ^ 1u);              // slot free (passes on first use)
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int it = pw + 4 * u;
                    const int g = it % 16, cq = it / 16;
                    const int row = g * 8 + (lane & 7), kf = (cq * 4 + (lane >> 3)) * 4;
                    float4 h, l;
                    um_split(v[u].x, h.x, l.x); um_split(v[u].y, h.y, l.y); um_split(v[u].z, h.z, l.z); um_split(v[u].w, h.w, l.w);
                    const unsigned o = um_off(row, kf);
                    *reinterpret_cast<float4 *>(dh + o) = h;
                    *reinterpret_cast<float4 *>(dl + o) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) um_mbar_arrive(b_full + 8 * stage);
                issue(c + (long long)DD * STAGES, slab);                          // refill the slab just consumed
                slab = slab + 1 == DD ? 0 : slab + 1;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer =====
        if (lane == 0) {
            unsigned it_count = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
                const unsigned buf = tcount & 1u, tuse = tcount >> 1;
                um_mbar_wait(b_tempty + 8 * buf, (tuse & 1u) ^ 1u);              // accumulator drained (passes on first use)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned acc = tmem + buf * (unsigned)N;
                for (int j = 0; j < nchunk; j++, it_count++) {
                    const int stage = it_count % STAGES;
                    const unsigned use = it_count / STAGES;
                    um_mbar_wait(b_full + 8 * stage, use & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int kc = min(UM_KC, K - j * UM_KC);
                    const unsigned ah = sa + stage * a_stage, al = ah + 128 * UM_KC * 4;
                    const unsigned bh = sw_hi + j * w_bytes, bl = sw_lo + j * w_bytes;
                    for (int s = 0; s < kc / 8; s++) {
                        const unsigned adv = (unsigned)s * 256;
                        const unsigned long long dah = um_desc(ah + adv, lbo, sbo), dal = um_desc(al + adv, lbo, sbo);
                        const unsigned long long dbh = um_desc(bh + adv, lbo, sbo), dbl = um_desc(bl + adv, lbo, sbo);
                        um_mma_tf32(acc, dah, dbh, idesc, (j > 0 || s > 0) ? 1u : 0u);
                        um_mma_tf32(acc, dal, dbh, idesc, 1u);
                        um_mma_tf32(acc, dah, dbl, idesc, 1u);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_empty + 8 * stage) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_tfull + 8 * buf) : "memory");
            }
        }
    } else {
        // ===== epilogue (warps 0-3) =====
        unsigned tcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
            const unsigned buf = tcount & 1u, tuse = tcount >> 1;
            um_mbar_wait(b_tfull + 8 * buf, tuse & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // 32-column panels: TMEM -> registers (lane = row) -> shared (row-major, 144-byte row pitch: conflict free) ->
            // registers (8 lanes = one 128-byte row segment) -> global: every store instruction writes 4 full 128-byte lines
            // instead of 32 scattered 16-byte pieces
            float *stg = reinterpret_cast<float *>(pstage) + warp * (32 * UM2_EPI_PITCH);
            const long long trow0 = (long long)tile * 128 + warp * 32;
            for (int col = 0; col < N; col += 32) {
                const int pw_cols = min(32, N - col);                 // 32 or 16
                for (int half = 0; half < pw_cols; half += 16) {
                    unsigned r[16];
                    const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + buf * (unsigned)N + (unsigned)(col + half);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                                 : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float4 *d4 = reinterpret_cast<float4 *>(stg + lane * UM2_EPI_PITCH + half);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        d4[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                            __uint_as_float(r[4 * q + 3]));
                }
                __syncwarp();
                const int q = lane & 7;                               // float4 within the 128-byte row segment
                if (4 * q < pw_cols) {
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias) bv = __ldg(reinterpret_cast<const float4 *>(bias + n0 + col) + q);
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        const int rr = it * 4 + (lane >> 3);
                        if (trow0 + rr < n) {
                            float4 o = *reinterpret_cast<const float4 *>(stg + rr * UM2_EPI_PITCH + 4 * q);
                            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                            *reinterpret_cast<float4 *>(Y + (size_t)(trow0 + rr) * ldy + n0 + col + 4 * q) = o;
                        }
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) um_mbar_arrive(b_tempty + 8 * buf);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)ncols));
}

static int g_umma_v = 3;    // kernel version used by cb_umma_linear: 3 = persistent pipelined, cp.async raw ring (default),
                            // 2 = persistent pipelined, loads held in registers, 1 = simple one-tile CTAs
extern "C" int cb_linear_set_umma_version(int v)
{
    if (v >= 1 && v <= 3) g_umma_v = v;
    return g_umma_v;
}
static size_t um_ring_bytes(int v)
{
    return v == 3 ? (size_t)UM3_STAGES * 2 * 128 * UM_KC * 4 + UM2_EPI_BYTES + (size_t)UM3_STAGES * UM3_DEPTH * 128 * UM_KC * 4
                  : (size_t)UM2_STAGES * 2 * 128 * UM_KC * 4 + UM2_EPI_BYTES;
}

static int g_umma = 1;      // 1: tcgen05 path for the shapes it takes (default) | 0: mma.sync kernels (tc_gemm.cu)
extern "C" int cb_linear_set_umma(int on)
{
    if (on == 0 || on == 1) g_umma = on;
    return g_umma;
}
int cb_umma_enabled() { return g_umma; }

bool cb_umma_shape_ok(int n, int K, int N, const float *A, const float *Y, int lda, int ldy)
{
    // W (TF32 hi + lo, all of K) of one COLUMN BLOCK stays resident in shared memory next to the A ring; the blocks are
    // gridDim.y of one launch.  Needs at least 16 columns per block: K <= 864.
    const int nchunk = (K + UM_KC - 1) / UM_KC;
    const size_t per_col = (size_t)2 * nchunk * UM_KC * 4;
    const size_t a_ring = um_ring_bytes(2);     // v3 falls back to the v2 ring when W does not fit beside its raw ring
    // 32 -> 32 layers: 12 MMAs of N = 32 per 32 KB tile are issue-latency bound (26.7 us vs 20.4 us for the mma.sync kernel at
    // n = 163840, tools/bench_linear.py); from 64 columns or 64 reduction elements on, the tcgen05 kernel is 1.1-1.6x faster
    return g_umma && n > 0 && K >= 8 && K % 8 == 0 && N >= 16 && N % 16 == 0 && (K >= 64 || N >= 64) && lda % 4 == 0 && ldy % 4 == 0 &&
           (((uintptr_t)A | (uintptr_t)Y) & 15) == 0 && 16 * per_col + a_ring + 2048 <= (size_t)227 * 1024;
}

// Y (n x N) = A (n x K) . B^T (+ bias); trans_b = 0: W is (N x K) (forward), 1: W is (K x N) (dgrad).  N is processed in
// column blocks of <= 256 (one TMEM allocation each).
int cb_umma_linear(int n, int K, int N, const float *A, int lda, const float *W, int trans_b, const float *bias, float *Y, int ldy,
                   cudaStream_t st)
{
    const int ldw = trans_b ? N : K;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    static bool attr_set = false;
    if (!attr_set) {
        // ONE opt-in to the full 227 KB for every instance: the dynamic size differs from layer to layer, and a CUDA graph
        // replays a node with ITS size against whatever the attribute is at replay time
        const int dyn_max = 227 * 1024 - 2048;        // 227 KB per CTA minus the kernels' static shared memory (1 KB) and slack
        cudaFuncSetAttribute(k_umma_linear2<0, UM2_STAGES, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max);
        cudaFuncSetAttribute(k_umma_linear2<1, UM2_STAGES, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max);
        cudaFuncSetAttribute(k_umma_linear2<0, UM3_STAGES, UM3_DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max);
        cudaFuncSetAttribute(k_umma_linear2<1, UM3_STAGES, UM3_DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max);
        cudaFuncSetAttribute(k_umma_linear<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max);
        cudaFuncSetAttribute(k_umma_linear<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max);
        (void)cudaGetLastError();
        attr_set = true;
    }
    if (g_umma_v >= 2) {
        // column blocks such that W (hi + lo, all of K) + the A ring fit the 227 KB of one CTA; N of a block % 16 == 0, <= 256
        const int nchunk = (K + UM_KC - 1) / UM_KC;
        int ver = g_umma_v;
        if (ver == 3) {                      // the raw ring costs 96 KB: only where all of W still fits in ONE column block
            const size_t room = (size_t)227 * 1024 - 2048 - um_ring_bytes(3);
            if ((size_t)2 * nchunk * N * UM_KC * 4 > room || N > 256) ver = 2;
        }
        const size_t a_ring = um_ring_bytes(ver);
        int nb_max = (int)(((size_t)227 * 1024 - 2048 - a_ring) / ((size_t)2 * nchunk * UM_KC * 4));
        nb_max = nb_max / 16 * 16;
        if (nb_max > 256) nb_max = 256;
        if (nb_max >= 16) {
            const int nblocks = (N + nb_max - 1) / nb_max;
            const int nb = ((N + nblocks - 1) / nblocks + 15) / 16 * 16;               // balanced column blocks
            const int ntiles = (n + 127) / 128;
            int ncols = 32;
            while (ncols < 2 * nb) ncols <<= 1;
            const size_t smem = (size_t)2 * nchunk * nb * UM_KC * 4 + a_ring;
            int bx = sms / nblocks;                                                    // persistent over row tiles, one wave
            if (bx < 1) bx = 1;
            if (bx > ntiles) bx = ntiles;
            const dim3 grid(bx, nblocks);
            if (ver == 3) {
                if (trans_b)
                    k_umma_linear2<1, UM3_STAGES, UM3_DEPTH><<<grid, UM3_THREADS, smem, st>>>(n, K, nb, ncols, ntiles, A, lda, W, bias, Y, ldy, N, ldw);
                else
                    k_umma_linear2<0, UM3_STAGES, UM3_DEPTH><<<grid, UM3_THREADS, smem, st>>>(n, K, nb, ncols, ntiles, A, lda, W, bias, Y, ldy, N, ldw);
            } else if (trans_b) {
                k_umma_linear2<1, UM2_STAGES, 0><<<grid, UM2_THREADS, smem, st>>>(n, K, nb, ncols, ntiles, A, lda, W, bias, Y, ldy, N, ldw);
            } else {
                k_umma_linear2<0, UM2_STAGES, 0><<<grid, UM2_THREADS, smem, st>>>(n, K, nb, ncols, ntiles, A, lda, W, bias, Y, ldy, N, ldw);
            }
            CB_COUNT(1);
            return CB_OK;
        }
    }
    for (int c0 = 0; c0 < N; c0 += 256) {
        const int nb = N - c0 < 256 ? N - c0 : 256;
        int ncols = 32;
        while (ncols < nb) ncols <<= 1;
        const size_t smem = (size_t)(2 * 128 + 2 * nb) * UM_KC * 4;
        const int blocks = (n + 127) / 128;
        if (trans_b) {
            k_umma_linear<1><<<blocks, UM_THREADS, smem, st>>>(n, K, nb, ncols, A, lda, W, bias, Y, ldy, c0, ldw);
        } else {
            k_umma_linear<0><<<blocks, UM_THREADS, smem, st>>>(n, K, nb, ncols, A, lda, W, bias, Y, ldy, c0, ldw);
        }
        CB_COUNT(1);
    }
    return CB_OK;
}

// stand-alone entry point (tests, microbench): see include/cbops.h
extern "C" int cb_umma_linear_forward(int n, int ci, int co, const float *X, const float *W, const float *b, float *Y, void *stream)
{
    CB_REQUIRE(n >= 0 && ci > 0 && co > 0 && X && W && Y, CB_EINVAL, "cb_umma_linear_forward: bad arguments");
    CB_REQUIRE(ci % 8 == 0 && co % 16 == 0 && ((((uintptr_t)X | (uintptr_t)Y) & 15) == 0), CB_EUNSUPPORTED,
               "cb_umma_linear_forward: needs ci %% 8 == 0, co %% 16 == 0, 16-byte aligned X / Y");
    if (n == 0) return CB_OK;
    cb_umma_linear(n, ci, co, X, ci, W, 0, b, Y, co, (cudaStream_t)stream);
    CB_CUDA_CHECK("cb_umma_linear_forward");
    return CB_OK;
}

extern "C" int cb_umma_linear_dgrad(int n, int ci, int co, const float *G, const float *W, float *dX, void *stream)
{
    CB_REQUIRE(n >= 0 && ci > 0 && co > 0 && G && W && dX, CB_EINVAL, "cb_umma_linear_dgrad: bad arguments");
    CB_REQUIRE(co % 8 == 0 && ci % 16 == 0 && ((((uintptr_t)G | (uintptr_t)dX) & 15) == 0), CB_EUNSUPPORTED,
               "cb_umma_linear_dgrad: needs co %% 8 == 0, ci %% 16 == 0, 16-byte aligned G / dX");
    if (n == 0) return CB_OK;
    cb_umma_linear(n, co, ci, G, co, W, 1, nullptr, dX, ci, (cudaStream_t)stream);      // dX = G (n x co) . W (co x ci)
    CB_CUDA_CHECK("cb_umma_linear_dgrad");
    return CB_OK;
}
