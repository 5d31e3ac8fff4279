// Production kernel of the Word2Bits training path for sm_100a: TMA-staged row rings.
//
// One CTA = one corpus shard (one reference thread, src/word2bits.cpp:363-516), split
// into roles:
//   * producer warp (last warp): the reference's control flow — learning-rate schedule
//     (:379-393), sentence builder + sub-sampling (:394-413), shard termination
//     (:414-423), window draw and negative draws (:428-460) — and, for every position,
//     one cp.async.bulk (TMA) copy per embedding row into shared-memory rings:
//     context rows of u into the u-ring, target rows of v (group by group) into the
//     v-ring.  It runs ahead of the arithmetic by as many rows as the rings hold, which
//     is what keeps tens of kB per SM in flight on a latency-bound gather.
//   * consumer warps: thread t owns columns [4t, 4t+4) of every row.  Context average
//     (:431-449) and error accumulator (:486-488) are private registers; the dim-D dot
//     (:461-471) is a warp-shuffle tree plus one shared-memory hop per target group.
//     Updates leave through the TMA as well: each warp overwrites its 512-byte column
//     chunk of a landed v row with g*context_avg (:489-491) and issues
//     cp.reduce.async.bulk.global.add.f32 from that slot; the accumulated error goes to
//     every context row of u (:494-503) the same way from a double-buffered staging row.
// Flow control: mbarriers (complete_tx) for "rows landed", monotonic shared counters for
// "slots free again" (published after cp.async.bulk.wait_group.read).
//
// Ordering semantics: rows of position p+1.. are fetched before position p's updates
// land, so a context row shared by neighbouring positions is read one or two updates
// stale; no update is ever lost (all scatters are atomic adds in L2).  This is the same
// class of staleness the reference's Hogwild threads have (SURVEY §7 "hard parts");
// DESIGN.md quantifies it and tests/test_gpu_parity.py holds it to the L3 bars.
#pragma once
#include "w2b_kernels.cuh"

namespace w2b {

constexpr int kND = 4;       // descriptor ring depth (positions in flight)
constexpr int kMaxGrp = 8;   // max target groups per position

struct RingDesc {
  int cw, nt, exit_flag, us0, vs0;
  float alpha;
  int center, b;
  int ctx[2 * W2B_MAX_WINDOW];
  int tg[W2B_MAX_NEGATIVE + 1];
};

struct RingCtl {
  unsigned long long ubar[kND];
  unsigned long long vbar[kND][kMaxGrp];
  volatile int prog[32];  // per consumer warp: positions finished
  volatile int vrel[32];  // per consumer warp: v slots released
  volatile int urel;      // u slots released
  float red[2][16][32];   // [buffer][target in group][warp] partial dots
  double loss_out;
};

// ------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_reduce_add(void *dst, const void *src_smem, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// Shared-memory carve-up (host and device agree through these helpers).
struct RingLayout {
  int rowb, nu, nv;
  size_t off_uring, off_vring, off_err, off_desc, off_sen, off_ctl, total;
};
__host__ __device__ inline RingLayout ring_layout(long long D, int nu, int nv) {
  RingLayout L;
  L.rowb = (int)(D * 4);
  L.nu = nu;
  L.nv = nv;
  size_t o = 0;
  L.off_uring = o; o += (size_t)nu * L.rowb;
  L.off_vring = o; o += (size_t)nv * L.rowb;
  L.off_err = o;   o += (size_t)2 * L.rowb;
  o = (o + 15) & ~(size_t)15;
  L.off_desc = o;  o += sizeof(RingDesc) * kND;
  L.off_sen = o;   o += sizeof(int) * kMaxS;
  o = (o + 15) & ~(size_t)15;
  L.off_ctl = o;   o += sizeof(RingCtl);
  L.total = o;
  return L;
}

template <int BM, int G>
__global__ void __launch_bounds__(512, 1) train_ring_kernel(TrainParams p, int nu, int nv) {
  extern __shared__ __align__(128) unsigned char smem[];
  const RingLayout L = ring_layout(p.D, nu, nv);
  unsigned char *uring = smem + L.off_uring;
  unsigned char *vring = smem + L.off_vring;
  unsigned char *errbuf = smem + L.off_err;
  RingDesc *desc = reinterpret_cast<RingDesc *>(smem + L.off_desc);
  int *s_sen = reinterpret_cast<int *>(smem + L.off_sen);
  RingCtl *ctl = reinterpret_cast<RingCtl *>(smem + L.off_ctl);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncw = (blockDim.x >> 5) - 1;  // consumer warps; the last warp is the producer
  const int nct = ncw * 32;
  ShardState *shp = p.shards + p.shard_base + blockIdx.x;
  if (shp->done) return;
  const unsigned rowb = (unsigned)L.rowb;

  if (tid == 0) {
    for (int i = 0; i < kND; ++i) {
      mbar_init(&ctl->ubar[i], 1);
      for (int g = 0; g < kMaxGrp; ++g) mbar_init(&ctl->vbar[i][g], 1);
    }
    ctl->urel = 0;
    ctl->loss_out = 0.0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    ctl->prog[tid] = 0;
    ctl->vrel[tid] = 0;
  }
  __syncthreads();

  if (warp == ncw) {
    // ============================================================== producer warp
    const ShardState sh = *shp;
    unsigned long long r = sh.rng;
    long long cursor = sh.cursor;
    long long wc = sh.word_count, last = sh.last_word_count;
    const long long wc0 = wc;
    int len = 0, sp = 0, status = 0, done = 0;
    long long iters = 0;
    unsigned long long n_pos = 0, n_ctx = 0, n_tgt = 0;
    int q = 0;             // positions enqueued
    const int ngmax = (p.negative + 1 + G - 1) / G;
    int u_alloc = 0, v_alloc = 0;
    for (;;) {
      if (wc - last > 10000) {  // :379-393
        if (lane == 0) {
          unsigned long long delta = (unsigned long long)(wc - last) * (unsigned long long)p.wca_scale;
          long long wca = (long long)(atomicAdd(p.wca, delta) + delta);
          float a = __fmul_rn(p.starting_alpha, __fsub_rn(1.f, __fdiv_rn((float)wca, p.alpha_denom)));
          if ((double)a < (double)p.starting_alpha * 0.0001) a = (float)((double)p.starting_alpha * 0.0001);
          *(volatile float *)p.alpha = a;
        }
        last = wc;
      }
      if (len == 0) {
        if (p.word_budget > 0 && wc - wc0 >= p.word_budget) break;
        unsigned long long r2 = r;
        long long c2 = cursor, w2 = wc;
        int l2 = 0;
        status = build_sentence(p, sh, lane, s_sen, r2, c2, w2, l2);
        __syncwarp();
        if (status == 2) break;  // slice exhausted mid-sentence: nothing committed
        r = r2; cursor = c2; wc = w2; len = l2; sp = 0;
      }
      if (status == 1 || wc > p.shard_word_limit) {  // :414-423
        if (lane == 0) atomicAdd(p.wca, (unsigned long long)(wc - last) * (unsigned long long)p.wca_scale);
        last = wc;
        done = 1;
        break;
      }
      ++iters;
      // descriptor slot must have been consumed by every consumer warp
      const int slot = q % kND;
      if (q >= kND || p.serial) {
        for (;;) {
          if (q - ctl->prog[0] < (p.serial ? 1 : kND)) break;
          __nanosleep(64);
        }
      }
      RingDesc *d = &desc[slot];
      r = make_position(p, lane, s_sen, len, sp, r, d);
      __syncwarp();
      const int cw = d->cw, nt = d->nt;
      ++sp;
      if (sp >= len) len = 0;  // :505-509
      if (cw == 0) continue;   // single-word or empty sentence: one window draw, nothing trained
      n_pos += 1; n_ctx += cw; n_tgt += nt;
      // ---- context rows -> u-ring
      for (;;) {
        if (u_alloc + cw - ctl->urel <= nu) break;
        __nanosleep(64);
      }
      if (lane == 0) {
        d->us0 = u_alloc % nu;
        d->vs0 = v_alloc % nv;
        d->alpha = *(volatile float *)p.alpha;
        d->exit_flag = 0;
        mbar_expect_tx(&ctl->ubar[slot], (unsigned)cw * rowb);
      }
      __syncwarp();
      for (int k = lane; k < cw; k += 32)
        bulk_load(uring + (size_t)((u_alloc + k) % nu) * rowb, p.u + (long long)d->ctx[k] * p.D, rowb,
                  &ctl->ubar[slot]);
      u_alloc += cw;
      // ---- target rows -> v-ring, group by group
      // (every group barrier of the slot is armed for every position, with 0 bytes when the
      //  position has fewer groups, so that all barriers of a slot stay on the same phase)
      for (int g0 = 0, gi = 0; gi < ngmax; g0 += G, ++gi) {
        const int ng = max(0, min(G, nt - g0));
        if (ng == 0) {
          if (lane == 0) mbar_expect_tx(&ctl->vbar[slot][gi], 0);
          continue;
        }
        for (;;) {
          if (v_alloc + ng - ctl->vrel[0] <= nv) break;
          __nanosleep(64);
        }
        if (lane == 0) mbar_expect_tx(&ctl->vbar[slot][gi], (unsigned)ng * rowb);
        __syncwarp();
        if (lane < ng)
          bulk_load(vring + (size_t)((v_alloc + lane) % nv) * rowb, p.v + (long long)d->tg[g0 + lane] * p.D, rowb,
                    &ctl->vbar[slot][gi]);
        v_alloc += ng;
      }
      ++q;
    }
    // tell the consumers to stop (after the descriptor slot is free)
    {
      const int slot = q % kND;
      if (q >= kND) {
        for (;;) {
          if (q - ctl->prog[0] < kND) break;
          __nanosleep(64);
        }
      }
      if (lane == 0) {
        desc[slot].exit_flag = 1;
        desc[slot].cw = 0;
        desc[slot].nt = 0;
        mbar_expect_tx(&ctl->ubar[slot], 0);
      }
    }
    if (lane == 0) {
      shp->rng = r;
      shp->cursor = cursor;
      shp->word_count = wc;
      shp->last_word_count = last;
      shp->done = done;
      shp->n_iter = sh.n_iter + (unsigned long long)iters;
      shp->n_pos = sh.n_pos + n_pos;
      shp->n_ctx = sh.n_ctx + n_ctx;
      shp->n_tgt = sh.n_tgt + n_tgt;
    }
  } else {
    // ============================================================== consumer warps
    const bool active = tid < p.ncol;
    const int colb = tid * 16;  // byte offset of this thread's float4 in a row
    QParams qp;
    qp.bits = p.bitlevel;
    qp.seg = (p.bitlevel >= 4) ? exp2f((float)(p.bitlevel - 1)) : 1.f;
    double loss = 0.0;
    int rb = 0;
    // issuer (tid 0) bookkeeping: the v group whose deltas are written but not yet issued,
    // and the slots committed to the TMA but not yet confirmed read
    int iss_ng = 0, iss_vs = 0, iss_t0 = 0;
    const RingDesc *iss_d = nullptr;
    int rel = 0, unconfirmed = 0;
    for (int q = 0;; ++q) {
      const int slot = q % kND;
      const unsigned par = (unsigned)((q / kND) & 1);
      mbar_wait(&ctl->ubar[slot], par);
      const RingDesc *d = &desc[slot];
      if (d->exit_flag) break;
      const int cw = d->cw, nt = d->nt, us0 = d->us0;
      int vs = d->vs0;
      const float alpha = d->alpha;
      // ---- context gather + quantize + average (:431-449)
      float avg[4] = {0.f, 0.f, 0.f, 0.f};
      if (active) {
        for (int k = 0; k < cw; ++k) {
          int s = us0 + k; if (s >= nu) s -= nu;
          const float4 x = *reinterpret_cast<const float4 *>(uring + (size_t)s * rowb + colb);
          avg[0] = __fadd_rn(avg[0], quant<BM>(x.x, qp));
          avg[1] = __fadd_rn(avg[1], quant<BM>(x.y, qp));
          avg[2] = __fadd_rn(avg[2], quant<BM>(x.z, qp));
          avg[3] = __fadd_rn(avg[3], quant<BM>(x.w, qp));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) avg[i] = __fdiv_rn(avg[i], (float)cw);
      }
      float err[4] = {0.f, 0.f, 0.f, 0.f};
      // ---- targets (:450-492), G rows per step
      for (int g0 = 0, gi = 0; g0 < nt; g0 += G, ++gi) {
        const int ng = min(G, nt - g0);
        mbar_wait(&ctl->vbar[slot][gi], par);
        float4 x[G];
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (k < ng && active) {
            int s = vs + k; if (s >= nv) s -= nv;
            x[k] = *reinterpret_cast<const float4 *>(vring + (size_t)s * rowb + colb);
          } else {
            x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (k < ng) {
            float pd = 0.f;
            if (active) {
              pd = fmaf(avg[0], quant<BM>(x[k].x, qp), pd);
              pd = fmaf(avg[1], quant<BM>(x[k].y, qp), pd);
              pd = fmaf(avg[2], quant<BM>(x[k].z, qp), pd);
              pd = fmaf(avg[3], quant<BM>(x[k].w, qp), pd);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pd += __shfl_xor_sync(kFull, pd, o);
            if (lane == 0) ctl->red[rb][k][warp] = pd;
          }
        }
        consumer_bar(nct);
        // Every consumer is past (a) the context phase of this position when g0 == 0 and
        // (b) its delta writes of the previous group: the issuer hands that group to the TMA.
        if (tid == 0) {
          if (g0 == 0) ctl->urel = ctl->urel + cw;
          if (iss_ng) {
            for (int k = 0; k < iss_ng; ++k) {
              int s = iss_vs + k; if (s >= nv) s -= nv;
              float *dst = p.v + (long long)iss_d->tg[iss_t0 + k] * p.D;
              if (p.plain_store) bulk_store(dst, vring + (size_t)s * rowb, rowb);
              else bulk_reduce_add(dst, vring + (size_t)s * rowb, rowb);
            }
            bulk_commit();
            bulk_wait_read<1>();  // all but the group just committed have left shared memory
            rel += unconfirmed;
            unconfirmed = iss_ng;
            iss_ng = 0;
            ctl->vrel[0] = rel;
          }
        }
        if (warp == 0 && lane < ng) {  // reported loss (:480-483), one lane per target
          float f = 0.f;
          for (int w = 0; w < ncw; ++w) f += ctl->red[rb][lane][w];
          float dp = (g0 + lane == 0) ? f : -f;
          loss += (double)logf(sigmoid_report(dp));
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (k < ng) {
            float f = 0.f;
            for (int w = 0; w < ncw; ++w) f += ctl->red[rb][k][w];  // fixed order: deterministic
            const float g = grad_scalar(f, (g0 + k == 0) ? 1 : 0, alpha, p.exptab);
            if (active) {
              err[0] = fmaf(g, quant<BM>(x[k].x, qp), err[0]);  // :487, old v
              err[1] = fmaf(g, quant<BM>(x[k].y, qp), err[1]);
              err[2] = fmaf(g, quant<BM>(x[k].z, qp), err[2]);
              err[3] = fmaf(g, quant<BM>(x[k].w, qp), err[3]);
              float4 dv = make_float4(g * avg[0], g * avg[1], g * avg[2], g * avg[3]);  // :490
              if (p.plain_store) {
                dv.x += x[k].x; dv.y += x[k].y; dv.z += x[k].z; dv.w += x[k].w;
              }
              int s = vs + k; if (s >= nv) s -= nv;
              *reinterpret_cast<float4 *>(vring + (size_t)s * rowb + colb) = dv;  // in place over the landed row
            }
          }
        }
        fence_async_smem();  // generic-proxy writes above -> visible to the TMA after the next barrier
        rb ^= 1;
        if (tid == 0) {
          iss_ng = ng;
          iss_vs = vs;
          iss_t0 = g0;
          iss_d = d;
        }
        vs += ng; if (vs >= nv) vs -= nv;
      }
      // ---- scatter the error to every context row (:494-503) from a staging row
      unsigned char *eb = errbuf + (size_t)(q & 1) * rowb;
      if (active) *reinterpret_cast<float4 *>(eb + colb) = make_float4(err[0], err[1], err[2], err[3]);
      fence_async_smem();
      consumer_bar(nct);
      if (tid == 0) {
        for (int k = 0; k < iss_ng; ++k) {  // last target group of the position
          int s = iss_vs + k; if (s >= nv) s -= nv;
          float *dst = p.v + (long long)iss_d->tg[iss_t0 + k] * p.D;
          if (p.plain_store) bulk_store(dst, vring + (size_t)s * rowb, rowb);
          else bulk_reduce_add(dst, vring + (size_t)s * rowb, rowb);
        }
        for (int k = 0; k < cw; ++k) bulk_reduce_add(p.u + (long long)d->ctx[k] * p.D, eb, rowb);
        bulk_commit();
        if (p.serial) {
          bulk_wait_all();  // debug: all updates of this position are in L2 before the next one is fetched
          rel += unconfirmed + iss_ng;
          unconfirmed = 0;
        } else {
          bulk_wait_read<1>();
          rel += unconfirmed;
          unconfirmed = iss_ng;
        }
        iss_ng = 0;
        ctl->vrel[0] = rel;
        __threadfence_block();
        ctl->prog[0] = q + 1;  // every consumer has read this descriptor (it passed the barrier above)
      }
    }
    if (tid == 0) bulk_wait_all();
    if (warp == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(kFull, loss, o);
      if (lane == 0) shp->loss = shp->loss + loss;
    }
  }
}

}  // namespace w2b
