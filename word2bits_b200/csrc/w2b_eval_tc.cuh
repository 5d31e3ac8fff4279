// Tensor-core candidate pass of the analogy evaluator (src/compute-accuracy.c:150-177) for sm_100a.
//
// scores = Q (nq x D) . M^T (D x words) is the one dense contraction in this repository.  The reference's
// arg-max must be reproduced exactly (first index wins ties, only positive scores count, the three query
// words are skipped), so the tensor cores are used as a FILTER, not as the scorer:
//   pass 1 (this file): TF32 tcgen05.mma on the fp32 operands as they are (TMA -> 128B-swizzled shared memory
//          -> UMMA, accumulators in TMEM), 128 questions x 256 words per CTA; the epilogue reads the
//          accumulators back with tcgen05.ld, keeps the question's best approximate score so far (atomicMax)
//          and appends every (question, word) whose approximate score lies within 2*eps of that running best
//          to a candidate list.  TF32 drops the low 13 mantissa bits of either operand, so
//          |approx - exact| <= 2^-9 * |vec| * |m| (+ accumulation order): a bound, not a guess.
//   pass 2 (w2b_eval.cu): candidates still within 2*eps of the FINAL best are re-scored in fp32 in the
//          reference's operation order; the arg-max over them is the reference's arg-max.
// Layout: both operands K-major with a row pitch of Dp = D rounded up to 32 floats (zero padding), so a k-block
// is one 128-byte swizzle row; rows beyond nq / words are zero-filled by the TMA unit (OOB fill).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace w2b {
namespace tc {

constexpr int BM = 128;        // questions per CTA (UMMA M)
constexpr int BN = 256;        // words per CTA (UMMA N) = TMEM columns
constexpr int BK = 32;         // floats per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;      // tf32: 8 elements (32 bytes) per instruction
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*alignment*/ + 256 /*barriers*/;
constexpr int THREADS = 192;   // warp 0: TMA, warp 1: MMA + TMEM, warps 2-5: epilogue

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred P1;\nTC_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra TC_DONE;\nbra TC_WAIT;\nTC_DONE:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int x, int y, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
// K-major operand tile in 128B-swizzled shared memory: rows of 128 bytes, groups of 8 rows 1024 bytes apart.
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3FFFF) >> 4);  // start address, 16-byte units, bits [0,14)
  d |= (unsigned long long)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (unsigned long long)(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 bytes
  d |= (unsigned long long)1 << 46;                        // descriptor version (sm_100)
  d |= (unsigned long long)2 << 61;                        // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::tf32: D = f32, A = B = tf32, both K-major, M = 128, N = 256
constexpr unsigned kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ unsigned ordered(float s) {  // monotone map float -> uint for atomicMax
  const unsigned b = __float_as_uint(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct Candidate { int q, c; float s; };

// gmax[q] = ordered() of the best approximate score of question q over its valid, non-query words (0: none is
// positive).  cand[0 .. *n_cand) = every (q, c, approximate score) that was within 2*qeps[q] of the question's
// running best when its tile was finished (a superset of what is within 2*qeps[q] of the final best); entries
// beyond cand_cap are dropped and *n_cand keeps counting (the caller checks for overflow).
__global__ void __launch_bounds__(THREADS, 2)
eval_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapM, const int *q3,
               const float *qeps, unsigned *gmax, Candidate *cand, unsigned long long *n_cand, unsigned long long cand_cap,
               int nq, int words, int Dp) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char *sA = smem, *sB = smem + STAGES * A_BYTES;
  unsigned long long *bars = (unsigned long long *)(smem + STAGES * (A_BYTES + B_BYTES));
  unsigned *tmem_slot = (unsigned *)(bars + 8);
  const unsigned full0 = s32(bars), empty0 = s32(bars + STAGES), done = s32(bars + 2 * STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nk = Dp / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { bar_init(full0 + 8 * s, 1); bar_init(empty0 + 8 * s, 1); }
    bar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: BN fp32 columns x 128 lanes for the accumulator tile
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer
      for (int k = 0; k < nk; ++k) {
        const int s = k % STAGES;
        if (k >= STAGES) bar_wait(empty0 + 8 * s, ((k / STAGES) - 1) & 1);
        bar_expect_tx(full0 + 8 * s, A_BYTES + B_BYTES);
        tma_load_2d(s32(sA + s * A_BYTES), &mapQ, k * BK, m0, full0 + 8 * s);
        tma_load_2d(s32(sB + s * B_BYTES), &mapM, k * BK, n0, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer: one thread on behalf of the CTA
      for (int k = 0; k < nk; ++k) {
        const int s = k % STAGES;
        bar_wait(full0 + 8 * s, (k / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned long long da = umma_desc(s32(sA + s * A_BYTES)), db = umma_desc(s32(sB + s * B_BYTES));
#pragma unroll
        for (int j = 0; j < BK / UMMA_K; ++j)  // advance 32 bytes (2 x 16-byte units) inside the swizzle row
          umma_tf32(tmem, da + 2 * j, db + 2 * j, (k | j) ? 1u : 0u);
        umma_commit(empty0 + 8 * s);  // frees the stage when these MMAs have read it
      }
      umma_commit(done);              // accumulator complete
    }
  } else {
    // ---- epilogue: warp w may read TMEM lanes 32*(w%4) .. +31; thread = one question row
    const int quad = warp & 3;
    const int row = quad * 32 + lane, q = m0 + row;
    const bool qok = q < nq;
    const int b1 = qok ? q3[q * 3] : -1, b2 = qok ? q3[q * 3 + 1] : -1, b3 = qok ? q3[q * 3 + 2] : -1;
    // lower bound for candidates: the question's best so far (any tile, any CTA) minus the error window
    float thr = 0.f;
    if (qok) {
      const unsigned g = *(volatile unsigned *)(gmax + q);
      thr = (g ? __uint_as_float(g & 0x7fffffffu) : 0.f) - 2.f * qeps[q];
    }
    bar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float best = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      unsigned r[32];
      const unsigned taddr = tmem + ((unsigned)(quad * 32) << 16) + (unsigned)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = n0 + c0 + j;
        const float s = __uint_as_float(r[j]);
        if (qok && c < words && c != b1 && c != b2 && c != b3 && s > 0.f && s >= thr) {
          if (s > best) {
            best = s;
            thr = fmaxf(thr, s - 2.f * qeps[q]);  // (a superset is fine: thr only ever rises)
          }
          const unsigned long long at = atomicAdd(n_cand, 1ull);
          if (at < cand_cap) cand[at] = Candidate{q, c, s};
        }
      }
    }
    if (qok && best > 0.f) atomicMax(gmax + q, ordered(best));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN) : "memory");
}

// ---- host: 2-D tensor maps over the padded row-major operands (rows x Dp floats), box = BK floats x box_rows
typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline bool make_map(CUtensorMap *map, const float *base, long long rows, long long Dp, int box_rows) {
  static encode_fn enc = nullptr;
  if (!enc) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
    enc = (encode_fn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)Dp, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)Dp * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace w2b
