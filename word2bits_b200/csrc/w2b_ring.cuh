// Production kernel of the Word2Bits training path for sm_100a: TMA-staged row rings.
//
// One CTA = one corpus shard (one reference thread, src/word2bits.cpp:363-516), split
// into warp roles that communicate through shared memory:
//   * sampler warp: the reference's control flow — learning-rate schedule (:379-393),
//     sentence builder + sub-sampling (:394-413), shard termination (:414-423), window
//     draw and negative draws (:428-460).  Emits one descriptor per trained position
//     (context ids, target ids).  The unigram-table lookups of position p+1 are issued
//     while position p's descriptor is being finished (the draws are replayable by LCG
//     jump-ahead, so nothing depends on the previous lookups).
//   * loader warp: for every descriptor, one cp.async.bulk (TMA) copy per embedding row
//     into shared-memory rings — context rows of u into the u-ring, target rows of v
//     (group by group) into the v-ring.  It runs ahead of the arithmetic by as many rows
//     as the rings hold, which keeps tens of kB per SM in flight on a latency-bound gather.
//   * consumer warps, two phases per position:
//       context phase  (:431-449) thread t owns float4 column t of the cw landed u rows:
//                      quantize, sum in row order, divide by cw, publish context_avg.
//       target phase   (:450-492) one WARP per landed v row (rows dealt round-robin, R = 2 rows
//                      in flight per warp): lane l covers float4 columns l, l+32, ... ; it
//                      reads the rows, quantizes, dots against context_avg (4 independent FMA
//                      chains per row + interleaved 5-step shuffle all-reduces), takes g from
//                      the expTable,
//                      accumulates g*quantize(v) into private error registers (:487),
//                      overwrites the row in place with g*context_avg (:490) and hands the
//                      slot back to the TMA: cp.reduce.async.bulk.global.add.f32, one
//                      4*D-byte atomic-add scatter per row.
//     The per-warp error partials are summed through shared memory into a staging row that
//     is scattered to every context row of u (:494-503) with the same bulk reduce.
// Flow control: mbarriers (complete_tx) for "rows landed"; per-slot release counters
// (bumped after cp.async.bulk.wait_group.read) for "slot free again"; monotonic counters
// for descriptors.
//
// Ordering semantics: rows of position p+1.. are fetched before position p's updates
// land, so a context row shared by neighbouring positions is read one or two updates
// stale; no update is ever lost (all scatters are atomic adds in L2).  This is the same
// class of staleness the reference's Hogwild threads have (SURVEY section 7 "hard parts");
// ring_serial=1 turns the prefetch off for parity work.  DESIGN.md quantifies the effect
// and tests/test_gpu_parity.py holds it to the L3 bars.
#pragma once
#include "w2b_kernels.cuh"
#include "w2b_ptx.cuh"

namespace w2b {

constexpr int kND = 8;      // descriptor ring depth (positions in flight)
constexpr int kMaxGrp = 8;  // max target groups per position

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
  volatile int desc_ready;  // descriptors published by the sampler
  volatile int prog;        // positions whose descriptor is no longer needed by the consumers
  volatile int urel;        // u slots released
  float sf[2][W2B_MAX_NEGATIVE + 1];  // +-f of every target of a position (reported loss, :480-483)
  double loss_out;
};

// Shared-memory carve-up (host and device agree through this helper).
struct RingLayout {
  int rowb, nu, nv;
  size_t off_uring, off_vring, off_err, off_avg, off_errp, off_rc, off_desc, off_sen, off_ctl, total;
};
__host__ __device__ inline RingLayout ring_layout(long long D, int nu, int nv, int ncw) {
  RingLayout L;
  L.rowb = (int)(D * 4);
  L.nu = nu;
  L.nv = nv;
  size_t o = 0;
  L.off_uring = o; o += (size_t)nu * L.rowb;
  L.off_vring = o; o += (size_t)nv * L.rowb;
  L.off_err = o;   o += (size_t)2 * L.rowb;        // staging rows for the u scatter (double-buffered)
  L.off_avg = o;   o += (size_t)L.rowb;            // context_avg
  L.off_errp = o;  o += (size_t)ncw * L.rowb;      // per-warp error partials
  L.off_rc = o;    o += sizeof(int) * (size_t)nv;  // per-slot release counters
  o = (o + 15) & ~(size_t)15;
  L.off_desc = o;  o += sizeof(RingDesc) * kND;
  L.off_sen = o;   o += sizeof(int) * kMaxS;
  o = (o + 15) & ~(size_t)15;
  L.off_ctl = o;   o += sizeof(RingCtl);
  L.total = o;
  return L;
}

// Division-free index arithmetic of the OPT = 1 variant: floor(x / d) == umulhi(x, ceil(2^32 / d)) exactly
// whenever x * d < 2^32 and d >= 2; d == 1 is encoded as magic == 0 (tests/test_host_cpu.py checks the
// identity exhaustively over the kernel's operand ranges through w2b_host_ring_index).
__host__ __device__ inline unsigned ring_magic(unsigned d) {
  return d <= 1u ? 0u : (unsigned)((0x100000000ull + d - 1u) / d);
}
__host__ __device__ inline unsigned ring_div(unsigned x, unsigned magic) {
#ifdef __CUDA_ARCH__
  return magic ? __umulhi(x, magic) : x;
#else
  return magic ? (unsigned)(((unsigned long long)x * magic) >> 32) : x;
#endif
}
// slot of target row i of a position whose first row sits in slot vs0 (< nv), and the landing barrier
// (group) the row belongs to
__host__ __device__ inline void ring_row_index(unsigned vs0, unsigned i, unsigned nv, unsigned nv_magic,
                                               unsigned g_magic, unsigned *slot, unsigned *group) {
  const unsigned x = vs0 + i;
  *slot = x - ring_div(x, nv_magic) * nv;
  *group = ring_div(i, g_magic);
}

// TrainParams::serial: 0 = production (prefetching), 1 = parity aid (no prefetch across positions).  The
// variants (OPT = 1) also know 2 = "early release": a unit leader confirms the previous pass's bulk reduces at
// the top of its next pass (they were issued a whole pass earlier, so the wait is free) and hands their slots
// back one pass sooner than the measured kernel, which releases them when the next pass commits — i.e., with
// two passes per position, only at the end of the position, so that most target rows of position p+1 cannot even
// be requested before position p is finished.  The measured kernel (OPT = 0) treats any non-zero value as 1.
template <int OPT>
__device__ __forceinline__ bool ring_serial(const TrainParams &p) { return OPT ? p.serial == 1 : p.serial != 0; }

// OPT = 0: the kernel measured in round 1.  OPT = 1 (cfg.kernel = 2): same protocol
// and arithmetic, fewer instructions on the consumer warps' critical path — slot / group indices without
// integer division (the signed / and % by run-time nv and G cost ~100 SASS instructions per 2-row batch in
// front of the first row load), compile-time column offsets for all but the last float4 column group, and
// running slot counters in the loader.  Kept as a variant until it has been measured on the GPU.
//
// LPR (lanes per row, cfg.kernel = 3: 16, cfg.kernel = 4: 8; always with OPT = 1) is the narrow-row variant:
// a warp is split into 32 / LPR row units that each own a target row of the batch, so one pass of the row
// loop serves 2 (4) x R rows — the per-batch fixed costs (barrier waits, index arithmetic, shuffle tree,
// expTable lookup, bulk-reduce issue) are shared by twice (four times) as many rows, the shuffle tree is one
// (two) steps shorter, and D = 400 fills 100 of 112 lane slots instead of 100 of 128.  NJ counts float4
// columns per lane: ceil(D / 4 / LPR).  Each unit's leader lane issues and confirms its own bulk reduces.
//
// XW (cfg.kernel = 5: 2, with OPT = 1) adds consumer warps beyond one per 128 columns: ncu shows the measured
// kernel stalled on fixed-latency dependencies with 2.25 warps per scheduler; at D = 800 the register file holds
// 11 warps of this kernel, and 25 target rows over 9 warps need a 2 + 1-row second pass instead of 2 + 2.
template <int BM, int NJ, int R, int OPT = 0, int LPR = 32, int XW = 0>
__global__ void __launch_bounds__((((NJ * LPR + 31) / 32 < 4 ? 4 : (NJ * LPR + 31) / 32) + 2 + XW) * 32, LPR == 32 ? 1 : 2)
    train_ring_kernel(TrainParams p, int nu, int nv, int G) {  // narrow-row variants: two CTAs per SM (<= 168 registers)
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncw = (blockDim.x >> 5) - 2;  // consumer warps; then the loader warp, then the sampler warp
  const int nct = ncw * 32;
  constexpr int UPW = 32 / LPR;  // row units per warp
  const RingLayout L = ring_layout(p.D, nu, nv, ncw);
  const unsigned s_base = smem_u32(smem);
  const unsigned uring = s_base + (unsigned)L.off_uring;
  const unsigned vring = s_base + (unsigned)L.off_vring;
  const unsigned errbuf = s_base + (unsigned)L.off_err;
  const unsigned s_avg = s_base + (unsigned)L.off_avg;
  const unsigned s_errp = s_base + (unsigned)L.off_errp;
  volatile int *s_rc = reinterpret_cast<volatile int *>(smem + L.off_rc);
  RingDesc *desc = reinterpret_cast<RingDesc *>(smem + L.off_desc);
  int *s_sen = reinterpret_cast<int *>(smem + L.off_sen);
  RingCtl *ctl = reinterpret_cast<RingCtl *>(smem + L.off_ctl);
  ShardState *shp = p.shards + p.shard_base + blockIdx.x;
  if (shp->done) return;
  const unsigned rowb = (unsigned)L.rowb;
  const int D4 = p.ncol;  // float4 columns per row
  const unsigned ubar0 = smem_u32(&ctl->ubar[0]);
  const unsigned vbar0 = smem_u32(&ctl->vbar[0][0]);

  if (tid == 0) {
    for (int i = 0; i < kND; ++i) {
      mbar_init(&ctl->ubar[i], 1);
      for (int g = 0; g < kMaxGrp; ++g) mbar_init(&ctl->vbar[i][g], 1);
    }
    ctl->desc_ready = 0;
    ctl->urel = 0;
    ctl->prog = 0;
    ctl->loss_out = 0.0;
#ifndef W2B_EMULATE
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  }
  for (int i = tid; i < nv; i += blockDim.x) s_rc[i] = 0;
  __syncthreads();

  if (warp == ncw + 1) {
    // ================================================================ sampler warp
    const ShardState sh = *shp;
    unsigned long long r = sh.rng;
    long long cursor = sh.cursor;
    long long wc = sh.word_count, last = sh.last_word_count;
    const long long wc0 = wc;
    int len = 0, sp = 0, status = 0, done = 0;
    long long iters = 0;
    unsigned long long n_pos = 0, n_ctx = 0, n_tgt = 0;
    int q = 0;  // descriptors published
    const int W = p.window, neg = p.negative;
    const int negl = neg < 32 ? neg : 32;
    const unsigned long long JA1 = c_JA[lane + 1], JC1 = c_JC[lane + 1];  // lane's own jump constants
    const unsigned long long JAn = c_JA[neg], JCn = c_JC[neg];
    bool have_pre = false;
    unsigned long long r1_pre = 0, rd_pre = 0;
    int t_pre = 0;
    float alpha_c = *(volatile float *)p.alpha;
    for (;;) {
      if (wc - last > 10000) {  // :379-393
        unsigned long long delta = (unsigned long long)(wc - last) * (unsigned long long)p.wca_scale;
        long long wca = 0;
        if (lane == 0) wca = (long long)(atomicAdd(p.wca, delta) + delta);
        wca = __shfl_sync(kFull, wca, 0);
        float a = __fmul_rn(p.starting_alpha, __fsub_rn(1.f, __fdiv_rn((float)wca, p.alpha_denom)));
        if ((double)a < (double)p.starting_alpha * 0.0001) a = (float)((double)p.starting_alpha * 0.0001);
        if (lane == 0) *(volatile float *)p.alpha = a;
        alpha_c = a;
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
        have_pre = false;
        // the shared learning rate (:53) is re-read once per sentence: other shards move it
        // every 10k words each, by ~1e-4 relative per update
        alpha_c = *(volatile float *)p.alpha;
      }
      if (status == 1 || wc > p.shard_word_limit) {  // :414-423
        if (lane == 0) atomicAdd(p.wca, (unsigned long long)(wc - last) * (unsigned long long)p.wca_scale);
        last = wc;
        done = 1;
        break;
      }
      ++iters;
      // ---- draws of this position (:428-460)
      unsigned long long r1, rd;
      int t = 0;
      if (have_pre) {
        r1 = r1_pre; rd = rd_pre; t = t_pre;
      } else {
        r1 = lcg(r);
        rd = r1 * JA1 + JC1;
        if (lane < negl) t = p.table[(rd >> 16) % (unsigned long long)W2B_TABLE_SIZE];
      }
      have_pre = false;
      const int b = mod_small(r1, (unsigned)W);
      // descriptor slot: free once every consumer is done with position q - kND
      const int slot = q % kND;
      while (q - ctl->prog >= (ring_serial<OPT>(p) ? 1 : kND)) __nanosleep(p.sleep_ns);
      RingDesc *d = &desc[slot];
      const int center = len ? s_sen[sp] : -1;
      int cw = 0;
      if (len) {
        for (int a0 = b; a0 < 2 * W + 1 - b; a0 += 32) {
          const int a = a0 + lane;
          const int qq = sp - W + a;
          const bool ok = (a < 2 * W + 1 - b) && (a != W) && qq >= 0 && qq < len;
          const unsigned m = __ballot_sync(kFull, ok);
          if (ok) d->ctx[cw + __popc(m & ((1u << lane) - 1))] = s_sen[qq];
          cw += __popc(m);
        }
      }
      int nt = 0;
      if (cw) {
        const unsigned long long r_after = r1 * JAn + JCn;
        if (sp + 1 < len) {  // next position of the sentence: its draws are already determined
          r1_pre = lcg(r_after);
          rd_pre = r1_pre * JA1 + JC1;
          t_pre = 0;
          if (lane < negl) t_pre = p.table[(rd_pre >> 16) % (unsigned long long)W2B_TABLE_SIZE];
          have_pre = true;
        }
        if (lane == 0) d->tg[0] = center;
        nt = 1;
        {
          bool ok = lane < negl;
          int tt = t;
          if (ok && tt == 0) tt = (int)(rd % (unsigned long long)(p.V - 1)) + 1;  // :457
          ok = ok && (tt != center);                                                 // :458
          const unsigned m = __ballot_sync(kFull, ok);
          if (ok) d->tg[nt + __popc(m & ((1u << lane) - 1))] = tt;
          nt += __popc(m);
        }
        for (int d0 = 33; d0 <= neg; d0 += 32) {  // negative > 32: remaining draws, not prefetched
          const int k = d0 + lane;
          bool ok = k <= neg;
          int tt = 0;
          if (ok) {
            const unsigned long long rd2 = lcg_jump(r1, k);
            tt = p.table[(rd2 >> 16) % (unsigned long long)W2B_TABLE_SIZE];
            if (tt == 0) tt = (int)(rd2 % (unsigned long long)(p.V - 1)) + 1;
            ok = (tt != center);
          }
          const unsigned m = __ballot_sync(kFull, ok);
          if (ok) d->tg[nt + __popc(m & ((1u << lane) - 1))] = tt;
          nt += __popc(m);
        }
        r = r_after;
      } else {
        r = r1;
      }
      ++sp;
      if (sp >= len) len = 0;  // :505-509
      if (p.trace) {  // parity hook: one record per window draw, exactly what the oracle's trace holds
        __syncwarp();
        if (lane == 0) {
          const unsigned long long k = (*p.trace_n)++;
          if ((long long)k < p.trace_cap) {
            w2b_trace_rec *tr = p.trace + k;
            tr->center = center; tr->b = b; tr->cw = cw; tr->ntargets = nt; tr->alpha = alpha_c;
            for (int i = 0; i < nt; ++i) tr->targets[i] = d->tg[i];
          }
        }
        __syncwarp();
      }
      if (p.max_iters >= 0 && iters >= p.max_iters) { if (cw) { n_pos += 1; n_ctx += cw; n_tgt += nt; } break; }
      if (cw == 0) continue;   // single-word or empty sentence: one window draw, nothing trained
      n_pos += 1; n_ctx += cw; n_tgt += nt;
      if (!p.train) continue;  // draws only: nothing is handed to the loader / consumers
      if (lane == 0) {
        d->center = center; d->b = b; d->cw = cw; d->nt = nt;
        d->alpha = alpha_c;
        d->exit_flag = 0;
      }
      __syncwarp();
      ++q;
      if (lane == 0) {
        __threadfence_block();
        ctl->desc_ready = q;
      }
    }
    // tell the loader (and through it the consumers) to stop
    while (q - ctl->prog >= kND) __nanosleep(p.sleep_ns);
    if (lane == 0) {
      RingDesc *d = &desc[q % kND];
      d->exit_flag = 1;
      d->cw = 0;
      d->nt = 0;
      __threadfence_block();
      ctl->desc_ready = q + 1;
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
  } else if (warp == ncw) {
    // ================================================================= loader warp
    const int ngmax = (p.negative + 1 + G - 1) / G;
    int u_alloc = 0;
    int v_alloc = 0;  // rows handed to the v-ring so far; row i lives in slot i % nv on its (i / nv)-th use
    int u_slot = 0, v_slot = 0, v_use = 0;  // OPT: u_alloc % nu, v_alloc % nv, v_alloc / nv kept incrementally
    for (int q = 0;; ++q) {
      while (ctl->desc_ready <= q) __nanosleep(p.sleep_ns);
      __threadfence_block();
      const int slot = q % kND;
      RingDesc *d = &desc[slot];
      const unsigned ubar = ubar0 + slot * 8;
      if (d->exit_flag) {
        if (lane == 0) mbar_expect_tx(ubar, 0);
        break;
      }
      const int cw = d->cw, nt = d->nt;
      // ---- context rows -> u-ring
      while (u_alloc + cw - ctl->urel > nu) __nanosleep(p.sleep_ns);
      if (lane == 0) {
        d->us0 = OPT ? u_slot : u_alloc % nu;
        d->vs0 = OPT ? v_slot : v_alloc % nv;
        mbar_expect_tx(ubar, (unsigned)cw * rowb);
      }
      __syncwarp();
      if constexpr (OPT) {
        for (int k = lane; k < cw; k += 32) {  // cw <= 2 * window <= nu: one wrap at most
          int us = u_slot + k;
          if (us >= nu) us -= nu;
          bulk_load(uring + (unsigned)us * rowb, p.u + (long long)d->ctx[k] * p.D, rowb, ubar);
        }
        u_slot += cw;
        if (u_slot >= nu) u_slot -= nu;
      } else {
        for (int k = lane; k < cw; k += 32)
          bulk_load(uring + (unsigned)((u_alloc + k) % nu) * rowb, p.u + (long long)d->ctx[k] * p.D, rowb, ubar);
      }
      u_alloc += cw;
      // ---- target rows -> v-ring, group by group.  Every group barrier of the slot is armed
      // for every position (0 bytes when the position has fewer groups) so that all barriers
      // of a descriptor slot stay on the same phase.
      for (int g0 = 0, gi = 0; gi < ngmax; g0 += G, ++gi) {
        const int ng = max(0, min(G, nt - g0));
        const unsigned vbar = vbar0 + (slot * kMaxGrp + gi) * 8;
        if (lane == 0) mbar_expect_tx(vbar, (unsigned)ng * rowb);
        __syncwarp();
        if (lane < ng) {  // each lane waits for its own slot to have been released by its last user
          int sl, uses;
          if constexpr (OPT) {  // ng <= G <= nv / 2: one wrap at most
            sl = v_slot + lane;
            uses = v_use;
            if (sl >= nv) { sl -= nv; ++uses; }
          } else {
            const int vi = v_alloc + lane;
            sl = vi % nv;
            uses = vi / nv;
          }
          while (s_rc[sl] < uses) __nanosleep(p.sleep_ns);
          bulk_load(vring + (unsigned)sl * rowb, p.v + (long long)d->tg[g0 + lane] * p.D, rowb, vbar);
        }
        __syncwarp();
        v_alloc += ng;
        if constexpr (OPT) {
          v_slot += ng;
          if (v_slot >= nv) { v_slot -= nv; ++v_use; }
        }
      }
    }
  } else {
    // ============================================================== consumer warps
    QParams qp;
    qp.bits = p.bitlevel;
    qp.seg = (p.bitlevel >= 4) ? exp2f((float)(p.bitlevel - 1)) : 1.f;
    double loss = 0.0;                 // warp 0: one lane per target
    // row units: LPR lanes that own one target row of a batch (the whole warp when LPR == 32)
    const int sub = (LPR == 32) ? 0 : lane / LPR;
    const int ul = (LPR == 32) ? lane : lane % LPR;  // lane within its unit
    const int nunits = ncw * UPW;
    const bool leader = ul == 0;       // issues and confirms the unit's bulk reduces
    int prev[R];                       // leader: slots whose reduce is committed but not yet confirmed read
#pragma unroll
    for (int t = 0; t < R; ++t) prev[t] = -1;
    const bool issuer = (warp == ncw - 1) && lane == 0;  // the last warp has the fewest rows: it scatters u
    const RingDesc *pend_u = nullptr;  // issuer: position whose u scatter is staged but not yet issued
    int pend_q = 0;
    // lane's float4 columns j*32+lane; columns past the row end are clamped for loads and
    // masked for stores (their context_avg registers are zero, so they add nothing)
    bool on[NJ];
    unsigned coff[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = j * LPR + ul;
      if (OPT && j < NJ - 1) {
        // NJ == ceil(D4 / LPR) (pick_ring): every column group but the last is full, so its lanes are always
        // on and its byte offset is ul * 16 + a compile-time constant (an LDS/STS immediate)
        on[j] = true;
        coff[j] = (unsigned)ul * 16u + (unsigned)j * (unsigned)(LPR * 16);
      } else {
        on[j] = c < D4;
        coff[j] = (unsigned)(on[j] ? c : D4 - 1) * 16u;
      }
    }
    const unsigned nv_magic = ring_magic((unsigned)nv), g_magic = ring_magic((unsigned)G);
    const bool col_on = tid < D4;
    const unsigned colb = (unsigned)(col_on ? tid : 0) * 16u;

    for (int q = 0;; ++q) {
      const int slot = q % kND;
      const unsigned par = (unsigned)((q / kND) & 1);
      mbar_wait(ubar0 + slot * 8, par);
      const RingDesc *d = &desc[slot];
      const bool fin = d->exit_flag != 0;
      const int cw = d->cw, nt = d->nt, us0 = d->us0, vs0 = d->vs0;
      const float alpha = d->alpha;
      // ---- context phase: gather + quantize + average (:431-449), thread per float4 column
      if (!fin && col_on) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int s = us0;
        for (int k = 0; k < cw; ++k) {
          const float4 x = lds128(uring + (unsigned)s * rowb + colb);
          a0 = __fadd_rn(a0, quant<BM>(x.x, qp));
          a1 = __fadd_rn(a1, quant<BM>(x.y, qp));
          a2 = __fadd_rn(a2, quant<BM>(x.z, qp));
          a3 = __fadd_rn(a3, quant<BM>(x.w, qp));
          if (++s == nu) s = 0;
        }
        const float fcw = (float)cw;
        sts128(s_avg + colb, make_float4(__fdiv_rn(a0, fcw), __fdiv_rn(a1, fcw), __fdiv_rn(a2, fcw), __fdiv_rn(a3, fcw)));
      }
      consumer_bar(nct);  // A: context_avg visible; u rows consumed; previous staging row complete
      if (issuer) {
        if (!fin) ctl->urel = ctl->urel + cw;
        if (pend_u) {  // scatter of the previous position's error to its context rows (:494-503)
          const unsigned eb = errbuf + (unsigned)(pend_q & 1) * rowb;
          for (int k = 0; k < pend_u->cw; ++k) bulk_reduce_add(p.u + (long long)pend_u->ctx[k] * p.D, eb, rowb);
          bulk_commit();
          pend_u = nullptr;
          __threadfence_block();
          ctl->prog = pend_q + 1;  // that descriptor may be recycled
        }
      }
      if (fin) break;
      // ---- target phase (:450-492): one warp per landed v row, R rows in flight per warp
      // context_avg columns of this lane: re-read from shared memory at each use (registers are
      // capped at 168/thread with 9+ warps per CTA; the R row buffers need them more)
      float4 e[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) e[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      float *sf = ctl->sf[q & 1];
      for (int i0w = warp * UPW; i0w < nt; i0w += R * nunits) {
        // rows i0, i0+nunits, ... of this unit; a missing row re-reads a landed row and gets g = 0
        const int i0 = i0w + sub;
        // LPR < 32: a unit may have no row at all in the warp's last batch; it then re-reads the row of the
        // warp's first unit (i0w < nt) and, like every lane, waits for that row's barrier before touching it
        const int ifall = (LPR == 32) ? i0 : (i0 < nt ? i0 : i0w);
        if constexpr (OPT) {
          if (p.serial == 2 && leader) {  // early release of the previous pass's slots (see ring_serial)
            bool any = false;
#pragma unroll
            for (int t = 0; t < R; ++t) any |= prev[t] >= 0;
            if (any) {
              bulk_wait_read<0>();
#pragma unroll
              for (int t = 0; t < R; ++t)
                if (prev[t] >= 0) { s_rc[prev[t]] = s_rc[prev[t]] + 1; prev[t] = -1; }
            }
          }
        }
        int sl[R];
        unsigned row[R];
        bool have[R];
#pragma unroll
        for (int t = 0; t < R; ++t) {
          const int i = i0 + t * nunits;
          have[t] = i < nt;
          const int ii = have[t] ? i : ifall;
          int s;
          if constexpr (OPT) {
            unsigned us, ug;
            ring_row_index((unsigned)vs0, (unsigned)ii, (unsigned)nv, nv_magic, g_magic, &us, &ug);
            if (have[t] || LPR < 32) mbar_wait(vbar0 + ((unsigned)slot * kMaxGrp + ug) * 8u, par);
            s = (int)us;
          } else {
            if (have[t]) mbar_wait(vbar0 + (slot * kMaxGrp + ii / G) * 8, par);
            s = (vs0 + ii) % nv;  // a position may be longer than the ring (1+negative > nv)
          }
          sl[t] = s;
          row[t] = vring + (unsigned)s * rowb;
        }
        float4 x[R][NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
          for (int t = 0; t < R; ++t) x[t][j] = lds128(row[t] + coff[j]);
        float f[R];
        {
          float d0[R], d1[R], d2[R], d3[R];
#pragma unroll
          for (int t = 0; t < R; ++t) d0[t] = d1[t] = d2[t] = d3[t] = 0.f;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            float4 aj = lds128(s_avg + coff[j]);
            if (!on[j]) aj = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < R; ++t) {
              x[t][j] = make_float4(quant<BM>(x[t][j].x, qp), quant<BM>(x[t][j].y, qp), quant<BM>(x[t][j].z, qp),
                                    quant<BM>(x[t][j].w, qp));
              d0[t] = fmaf(aj.x, x[t][j].x, d0[t]);
              d1[t] = fmaf(aj.y, x[t][j].y, d1[t]);
              d2[t] = fmaf(aj.z, x[t][j].z, d2[t]);
              d3[t] = fmaf(aj.w, x[t][j].w, d3[t]);
            }
          }
#pragma unroll
          for (int t = 0; t < R; ++t) f[t] = (d0[t] + d1[t]) + (d2[t] + d3[t]);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)  // R interleaved butterflies: every lane of a unit ends with the full sums
#pragma unroll
          for (int t = 0; t < R; ++t) f[t] += __shfl_xor_sync(kFull, f[t], o);
        float g[R];
#pragma unroll
        for (int t = 0; t < R; ++t) {
          const int i = i0 + t * nunits;
          g[t] = have[t] ? grad_scalar(f[t], i == 0 ? 1 : 0, alpha, p.exptab) : 0.f;
          if (leader && have[t]) sf[i] = (i == 0) ? f[t] : -f[t];
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float4 aj = lds128(s_avg + coff[j]);
#pragma unroll
          for (int t = 0; t < R; ++t) {
            e[j].x = fmaf(g[t], x[t][j].x, e[j].x);  // :487, quantized OLD v
            e[j].y = fmaf(g[t], x[t][j].y, e[j].y);
            e[j].z = fmaf(g[t], x[t][j].z, e[j].z);
            e[j].w = fmaf(g[t], x[t][j].w, e[j].w);
            if (on[j] && have[t])  // :490 — the update g*context_avg replaces the landed row in its slot
              sts128(row[t] + coff[j], make_float4(g[t] * aj.x, g[t] * aj.y, g[t] * aj.z, g[t] * aj.w));
          }
        }
        fence_async_smem();
        __syncwarp();
        if (leader) {
#pragma unroll
          for (int t = 0; t < R; ++t)
            if (have[t]) bulk_reduce_add(p.v + (long long)d->tg[i0 + t * nunits] * p.D, row[t], rowb);
          bulk_commit();
          if (ring_serial<OPT>(p)) bulk_wait_all(); else bulk_wait_read<1>();
          // everything this lane committed before the group above has left shared memory
#pragma unroll
          for (int t = 0; t < R; ++t) {
            if (prev[t] >= 0) s_rc[prev[t]] = s_rc[prev[t]] + 1;
            prev[t] = have[t] ? sl[t] : -1;
            if (ring_serial<OPT>(p) && have[t]) { s_rc[sl[t]] = s_rc[sl[t]] + 1; prev[t] = -1; }
          }
        }
        __syncwarp();
      }
      if (leader) {  // confirm this unit's last rows (and, for the issuer, the u scatter) right away
        bulk_wait_read<0>();
#pragma unroll
        for (int t = 0; t < R; ++t) {
          if (prev[t] >= 0) s_rc[prev[t]] = s_rc[prev[t]] + 1;
          prev[t] = -1;
        }
      }
      // ---- error partials -> staging row (its scatter, :494-503, is issued after the next barrier A)
      if constexpr (LPR < 32) {  // the row units of a warp hold partials of the same columns: add them up first
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            e[j].x += __shfl_xor_sync(kFull, e[j].x, o);
            e[j].y += __shfl_xor_sync(kFull, e[j].y, o);
            e[j].z += __shfl_xor_sync(kFull, e[j].z, o);
            e[j].w += __shfl_xor_sync(kFull, e[j].w, o);
          }
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (on[j] && (LPR == 32 || sub == 0)) sts128(s_errp + (unsigned)warp * rowb + coff[j], e[j]);
      consumer_bar(nct);  // B
      if (col_on) {
        float4 acc = lds128(s_errp + colb);
        for (int w = 1; w < ncw; ++w) {
          const float4 t = lds128(s_errp + (unsigned)w * rowb + colb);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        sts128(errbuf + (unsigned)(q & 1) * rowb + colb, acc);
      }
      fence_async_smem();
      if (warp == 0)  // reported loss (:480-483): one lane per target, off the row loop's critical path
        for (int k = lane; k < nt; k += 32) loss += (double)logf(sigmoid_report(sf[k]));
      if (issuer) { pend_u = d; pend_q = q; }
      if (ring_serial<OPT>(p)) {  // parity aid: everything of this position lands before the next one is fetched
        consumer_bar(nct);
        if (issuer) {
          const unsigned eb = errbuf + (unsigned)(q & 1) * rowb;
          for (int k = 0; k < cw; ++k) bulk_reduce_add(p.u + (long long)d->ctx[k] * p.D, eb, rowb);
          bulk_commit();
          bulk_wait_all();
          pend_u = nullptr;
          __threadfence_block();
          ctl->prog = q + 1;
        }
      }
    }
    if (leader) bulk_wait_all();
    if (warp == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(kFull, loss, o);
      if (lane == 0) shp->loss = shp->loss + loss;
    }
  }
}

}  // namespace w2b
