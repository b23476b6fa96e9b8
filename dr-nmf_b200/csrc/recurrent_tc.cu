// The DR-NMF recurrence (custom_layers.py:343-375 under Keras' masked scan) as ONE persistent sm_100a kernel.
//
// Problem shape: per utterance the chain is T x K_layers strictly serial steps (layer 0 of frame t+1 starts from the
// last layer of frame t); per step the work is  g^k = relu(g^{k-1} . S_k + x~W_k + b_k + leak),  a (B x R).(R x R)
// product with a SKINNY batch dimension.  S_k (4 MB at R=1000) must be re-streamed every step, so many SMs have to
// pull a share of the weights each step: the step is tiled as
//       M-tile m (128 output atoms)  x  K-split s (a slice of Rp/KS input atoms)  x  batch group g,
//       grid = (KS, MT, G), cluster = the KS K-splits of one M-tile.
// What bounds a step (measured, profiles/r2_*): the split-K exchange inside the cluster moves a 128 x NB fp32 partial
// tile per CTA through distributed shared memory at ~13 B/clk (700 + 40 NB cycles), the hop to the consumers through
// L2 costs ~3.5k cycles whatever the bytes, while the 3xTF32 product of a 128-atom slice takes 26 NB cycles.  Hence
//   * throughput regime (B > 64): K-splits of 4 (cluster of 4 CTAs: 33 clusters are co-resident on a B200, only 15
//     clusters of 8) - a CTA multiplies a 128 x (Rp/4) block, twice the tensor work per exchanged byte - and batch
//     groups: utterances are independent, so the batch is cut into G = (co-resident clusters) / MT groups that run the
//     whole chain on disjoint SMs with their own flags (R = 1000: 4 groups x 32 CTAs = 128 SMs), every group
//     pipelining 64-column batch tiles through the same weights;
//   * latency regime (B <= 64): the chain of dependent hops decides, not the bytes: K-splits of 8, one group, one or two
//     32-column tiles (see choose_plan).
//   * weights: the CTA's block of S_k^T - I streams through a shared-memory ring (TMA, 16 KB pieces of 32 K-columns),
//     is split into tf32 hi + remainder lo by the loader warps and lands in TENSOR MEMORY in 64-column sub-chunks
//     (3 buffers of hi|lo = 384 TMEM columns).  The MMA takes A from TMEM (no per-instruction smem read of A).  With
//     more sub-chunks than buffers (K-slice >= 256) consecutive batch tiles walk the K-slice in alternating direction,
//     so the sub-chunks resident at the turn are reused and only the rest is re-streamed (slot = sub-chunk mod 3);
//     the schedule (sub-chunk, slot, first use, last use per position) is built on the host and passed as a table.
//   * hidden state: B operand, 64-atom x NB sub-chunk tiles (hi and lo) TMA-loaded with 128B swizzle from a
//     ping-pong global buffer into a ring of stages; consumers acquire the producers' flags lazily per M-tile.
//   * product: tcgen05.mma kind::tf32, 3xTF32 compensation (W_lo.h_hi + W_hi.h_lo + W_hi.h_hi), fp32 accumulators in
//     TMEM, bursts of 24 back-to-back MMAs per sub-chunk from constant-offset descriptors.
//   * split-K reduction INSIDE the cluster over distributed shared memory (staging + cp.async.bulk + mbarrier
//     complete_tx): CTA o owns rows [o*RO, (o+1)*RO) of the M-tile, sums the KS partials in a fixed order
//     (deterministic), applies the fused epilogue (identity part, input projection, bias, rank-1 leak, relu, Keras mask
//     carry) and writes hi/lo of the new hidden rows; every owner warp releases its own stores with red.release.gpu
//     (measured faster at 1, 2 and 8 tiles per group than a publisher thread that batches the gpu-scope fences, which
//     remains selectable with DRNMF_REC_PUB=thread).
//   * latency regime, one tile per group: the hidden rows validate themselves (mantissa-LSB tag, see ll_tag) and two
//     consumer warps pull them with L2 loads - no flag, no release fence on the producer side.
//
// Warp roles (512 threads): 0 weight producer (TMA ring) | 1 hidden-state TMA (+flag acquire) | 2 MMA issuer / TMEM
//     owner | 3 publisher (or second consumer in the latency regime) | 4-7 TMEM -> DSMEM pushers | 8-11 row owners (reduce + epilogue) | 12-15 weight loaders
//     (smem ring -> regs -> TMEM).
#include "internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace drnmf {

constexpr int RT_THREADS = 512;
constexpr int RT_AST = 2;                    // TMEM accumulator stages
constexpr int RT_WB = 3;                     // TMEM weight buffers: 64 K-columns hi | 64 lo = 128 TMEM columns each
constexpr int RT_PST = 4;                    // owner -> publisher hand-off slots
constexpr int RT_MAXSC = 16;                 // sub-chunks (64 K-columns) per K-slice: K-slices of up to 1024 atoms
constexpr int RT_MAXH = 8;                   // hidden-state ring depth
constexpr int RT_MAXW = 16;                  // weight ring depth (16 KB pieces)
constexpr long long RT_WATCHDOG = 3000000000LL;

// Per-step schedule of a CTA: which 64-column sub-chunk of the K-slice each (batch tile, position) multiplies, in
// which TMEM weight buffer it lives, whether that buffer is filled for this use (first use) and whether it may be
// overwritten afterwards (last use).  Tile classes: 0 first tile, 1 odd tile, 2 even tile > 0; +3 when the tile is the
// last of the step.  Entry bits: [3:0] sub-chunk, [5:4] buffer, 6 first use, 7 last use.
// e_last: the same table for the last batch group when it holds fewer tiles (the classes depend on the tile count).
struct RecSched { uint8_t e[6][RT_MAXSC]; uint8_t e_last[6][RT_MAXSC]; };
using SchedTab = const uint8_t (*)[RT_MAXSC];

struct RecArgs {
  // tensors
  const float* XW; const float* mvalid; const float* h0;   // XW includes the bias b_k
  float *state, *psum, *Hp_hi, *Hp_lo, *H_user, *hb_hi, *hb_lo;
  float *actT_hi, *actT_lo;                  // training: K x Rp x (T*Bp) activations, time-major frames; else null
  // backward chain (BWD instantiation): dL/dH from the head, dL/dz^k out, state-gradient carry, partial row sums
  const float* dH; float *deltaT_hi, *deltaT_lo, *G, *psum2;
  const float* alph_vec;                     // K x Rp step sizes for vector alph (untie_alph) in the backward chain, else null
  float d0mo_b, o0_b, ok_b;
  unsigned int* flags;
  // pipelined forward: XW rows are time-major (t*Btot + b) and the frames >= xw_t0 are only valid once *xw_ready != 0
  const unsigned int* xw_ready; int xw_tmajor, xw_t0, Btot;
  unsigned int* started;                     // optional: every CTA counts itself in once it is resident (after the cluster sync)
  unsigned int* progress;                    // optional (backward chain, one tile, one group): completely processed frames
  int* dev_error;
  int dbg_m;                                 // M-tile of the observed CTA (K-split 0)
  long long* dbg;                            // optional per-role wait-time counters of CTA (0,0) (DRNMF_REC_DEBUG=1)
  long long* trace;                          // optional event trace of CTA (0,0): [0] = count, then (tag, clock) pairs
  int trace_lo, trace_hi;                    // items [lo, hi) of the observed CTA are traced
  int h3d;                                   // hidden-state tensor maps are 3-D slab maps (two 32-atom tiles per TMA instruction)
  int ll;                                    // latency mode: self-validating hidden-state exchange (see the consumer warp)
  int sym;                                   // S_k symmetric (scalar alph): sub-chunks below the diagonal are fetched mirrored
  // shapes
  int B, Bp, T, K, R, Rp;
  int MT, KS, RO, KSLICE, n_tiles;           // n_tiles = batch tiles per batch group (grid.z groups run on disjoint SMs)
  int n_tiles_total;                         // batch tiles of the whole batch
  int NSC;                                   // 64-column sub-chunks per K-slice (the last one may be 32 wide)
  int rot;                                   // NSC <= RT_WB: the buffers rotate by NSC per step (next step's weights prefetch)
  int split;                                 // forward, RO x NB = 4096: warps 4-7 push an item and then own the upper half of its rows
  int pub_unit;                              // flag increments per (CTA, item): 1 = publisher thread, 4 = each owner warp releases its own stores
  int WST, HST, RST;                         // weight-piece / hidden-sub-chunk / reduction-slot ring depths
  float u0_dmo, u0_off, uk_dmo, uk_off;
  // smem offsets (bytes from the 1024-aligned base)
  int off_w, off_h, off_red, off_push, off_leak, off_out, off_bar;
  int h_stage_bytes, red_slot_bytes;
};

struct RecBars {   // all mbarriers, laid out at off_bar
  uint64_t h_full[RT_MAXH], h_empty[RT_MAXH], t_full[RT_AST], t_empty[RT_AST], red_full[4], red_free[4];
  uint64_t pub_full[RT_PST], pub_empty[RT_PST];
  uint64_t wb_full[RT_WB], wb_empty[RT_WB];  // TMEM weight buffer holds its sub-chunk / may be overwritten
  uint64_t w_full[RT_MAXW], w_free[RT_MAXW]; // 16 KB weight pieces (32 K-columns x 128 rows) staged in smem by TMA
  uint32_t tmem_slot;
  int abort;
  long long dbg_ts[2];                       // debug: clock of the first satisfied h_full / of the last MMA issue of an item
  // debug accumulators live here, not in registers: 14 registers per thread of a kernel capped at 128 spilled loop
  // invariants of the production path (ptxas -v: ~500 bytes at NB = 32)
  long long dbg_acc[16][8];                  // [warp][0..6 = slots, 7 = start clock]
  int dbg_trc[8];                            // per-role trace counters
};

// Failure record (slow path, out of line): word 0 = first code, word 1 = mask of every code seen (bit = code - 200),
// so that a watchdog report tells which waits were stuck, not only which role started waiting first.
__device__ __noinline__ void rt_fail(int* e, int code) {
  atomicCAS(e, 0, code);
  if (code >= 200 && code < 232) atomicOr(e + 1, 1 << (code - 200));
}

// named-barrier AND-reduction over the 128 owner threads (barrier id 1): uniform agreement on a predicate
__device__ __forceinline__ bool owners_all(bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.and.pred p, 1, 128, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(r) : "r"((uint32_t)pred) : "memory");
  return r != 0;
}

__device__ __forceinline__ bool poll_flag(const unsigned int* f, unsigned int target, volatile int* err, long long budget = RT_WATCHDOG) {
  if (flag_ld_acquire(f) >= target) return true;
  long long t0 = clock64();
  unsigned it = 0;
  while (flag_ld_acquire(f) < target) {
    if ((++it & 0xFF) == 0) {
      if (clock64() - t0 > budget) return false;
      if (*err) return false;
    }
  }
  return true;
}

// accumulate the cycles spent in `stmt` into debug slot `slot` (only CTA (0,0), only when a.dbg is set; dbg_on is
// warp-uniform so that `stmt` may contain warp collectives, only the role's reporting lane accumulates)
#define DBG_ADD(slot, v) do { if (dbg_w) bars->dbg_acc[dbg_wi][slot] += (v); } while (0)
#define RT_TIMED(slot, stmt)                                   \
  do {                                                         \
    if (dbg_on) { long long _t0 = clock64(); stmt; DBG_ADD(slot, clock64() - _t0); } \
    else { stmt; }                                             \
  } while (0)

// event trace (debug): every tracing thread appends (tag, clock) pairs to its role's region of a.trace with a private
// counter (plain stores, no atomics: the probe must not move what it measures).  tag = event << 16 | item (low 16 bits)
constexpr int RT_TRC_PER_ROLE = 256;
#define RT_TRACE(role, ev, item)                                                                    \
  do {                                                                                              \
    if (dbg_on && a.trace && (long long)(item) >= a.trace_lo && (long long)(item) < a.trace_hi && bars->dbg_trc[role] < RT_TRC_PER_ROLE) { \
      const int _c = bars->dbg_trc[role];                                                           \
      long long* _p = a.trace + 8 + ((role) * RT_TRC_PER_ROLE + _c) * 2;                            \
      _p[0] = ((long long)(ev) << 16) | (long long)((item) & 0xFFFF); _p[1] = clock64();            \
      bars->dbg_trc[role] = _c + 1; a.trace[role] = _c + 1;                                         \
    }                                                                                               \
  } while (0)

// Latency mode ("LL" exchange, one batch tile per group): the new hidden rows carry their own validity.  Every fp32
// the owners write into the ping-pong buffer has its mantissa LSB forced to the parity of the write count of that
// slot (a <= 1 ulp perturbation that is part of the value from then on: the identity path re-reads the same bits), so a
// consumer needs no flag, no release fence on the producer side and no second round trip: it loads the slice, checks the
// LSB of every element against the parity it expects and repeats until all match.  Writes to slot (k & 1) happen at
// k = 0 .. K-2 of every frame: K/2 per frame to slot 0, (K-1)/2 to slot 1.
__device__ __forceinline__ uint32_t ll_tag(int frame, int k, int K) {
  const int per_frame = (k & 1) ? (K - 1) / 2 : K / 2;
  return (uint32_t)(frame * per_frame + (k >> 1) + 1) & 1u;      // +1: the first write differs from the zero-filled buffer
}
__device__ __forceinline__ float ll_mark(float x, uint32_t tag) { return __uint_as_float((__float_as_uint(x) & ~1u) | tag); }

// ring position without runtime division (a 64-bit % and / per use cost ~100 cycles each on the critical path)
struct RtRing {
  int idx; uint32_t ph;
  __device__ __forceinline__ RtRing() : idx(0), ph(0) {}
  __device__ __forceinline__ void next(int depth) { if (++idx == depth) { idx = 0; ph ^= 1u; } }
};

// schedule entry of (tile i of n, position p)
__device__ __forceinline__ uint32_t sched_at(SchedTab sc, int i, int n, int p) {
  const int cls = (i == 0 ? 0 : ((i & 1) ? 1 : 2)) + (i == n - 1 ? 3 : 0);
  return sc[cls][p];
}

// One pusher step (see the pusher role in the kernel) as a function: accumulator stage `as` -> staging slot `rs` -> bulk
// DSMEM copies into the owners' reduction slots.  Used by the split epilogue, where warps 4-7 push an item and then own
// half of its rows.  Executed by all 128 threads of warps 4-7.
template <int NB>
__device__ __forceinline__ void push_item(const RecArgs& a, RecBars* bars, uint8_t* smem, uint32_t tmem_base, uint32_t acc_col0,
                                          int s, int q, int lane, int as, uint32_t as_ph, int rs, uint32_t rs_ph, bool& dead) {
  constexpr int CHUNKS = NB / 4;
  constexpr int SWZ = (CHUNKS >= 8) ? 7 : CHUNKS - 1;
  volatile int* err = a.dev_error;
  const int rho = q * 32 + lane;
  const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t stage0 = smem_u32(smem + a.off_push);
  const uint32_t blk_bytes = (uint32_t)(a.RO * NB * 4);
  if (!dead && !mbar_wait(&bars->t_full[as], as_ph, err, RT_WATCHDOG)) { rt_fail(a.dev_error, 207); dead = true; }
  tc_fence_after();
  float v[NB];
#pragma unroll
  for (int c = 0; c < NB; c += 16) tmem_ld16(trow + acc_col0 + as * NB + c, v + c);
  tc_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (lane == 0 && !dead) mbar_arrive(&bars->t_empty[as]);
  if (!dead && !mbar_wait_cluster(&bars->red_free[rs], rs_ph ^ 1u, err, RT_WATCHDOG)) { rt_fail(a.dev_error, 208); dead = true; }
  const uint32_t srow = stage0 + rs * a.red_slot_bytes + rho * (NB * 4);
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const uint32_t addr = srow + (uint32_t)(((c & ~SWZ) | ((c ^ rho) & SWZ)) * 16);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                 "f"(v[4 * c + 2]), "f"(v[4 * c + 3]) : "memory");
  }
  fence_proxy_async_smem();
  int o_first, o_cnt;
  if (a.RO <= 32) { __syncwarp(); o_cnt = 32 / a.RO; o_first = q * o_cnt; }
  else { asm volatile("bar.sync 2, 128;" ::: "memory"); o_cnt = (q == 0) ? a.KS : 0; o_first = 0; }
  if (lane < o_cnt && !dead) {
    const uint32_t o = (uint32_t)(o_first + lane);
    const uint32_t src = stage0 + rs * a.red_slot_bytes + o * blk_bytes;
    const uint32_t dst = mapa_u32(smem_u32(smem + a.off_red) + rs * a.red_slot_bytes + (uint32_t)s * blk_bytes, o);
    const uint32_t bar = mapa_u32(smem_u32(&bars->red_full[rs]), o);
    dsmem_bulk_copy(dst, src, blk_bytes, bar);
  }
}

// Latency-mode consumer (warps 1 and 3, see ll_tag): warp `ci` serves the positions p = ci, ci + 2, ... of every step.
// Per 64-atom sub-chunk: NB rows x 16 float4; lane l takes float4 (l & 15) of rows (l >> 4) + 2 j.  The raw values go to
// the "hi" tiles, their tf32 remainders to the "lo" tiles, both in the 128B-swizzle layout of the MMA descriptors.  With
// NB = 16 the loads of the warp's next position are issued before the current one is stored (two register sets).
template <int NB>
__device__ __forceinline__ void ll_consumer(const RecArgs& a, SchedTab sch, RecBars* bars, uint8_t* smem, int ci, int s,
                                            int lane, bool dbg_on, bool dbg_w, int dbg_wi) {
  constexpr int NLD = NB / 2;                    // float4 per lane and sub-chunk
  constexpr bool PREFETCH = (NB == 16);
  constexpr uint32_t HB = NB * 128;
  volatile int* err = a.dev_error;
  const int K = a.K, T = a.T, Rp = a.Rp, NSC = a.NSC;
  const int q = lane & 15, r0 = lane >> 4;       // float4 index inside the sub-chunk, first row
  RtRing hr;
  long long item = 0;
  bool okh = true;
  // Polling is light: a consumer that spins on the whole slice keeps eight L2 readers on every line the owners are about
  // to write (measured: the stores queue behind the reads and the step stretches by 30 %).  The warp spins on the first
  // row of its sub-chunk only (16 lanes x 16 bytes) and pulls the full tile once that row is valid; every element is
  // still validated, stragglers are re-loaded.
  float4 v[NLD], w[NLD];
  const int n_tiles = a.n_tiles;
  auto issue = [&](float4 (&dst)[NLD], int slot, int i, int p) {
    const int sc = (int)(sched_at(sch, i, n_tiles, p) & 15u);
    const bool two = (a.KSLICE - sc * 64 >= 64);
    if (two || q < 8) {
      const float* src = a.hb_hi + ((size_t)slot * a.Bp + i * NB + r0) * Rp + s * a.KSLICE + sc * 64 + 4 * q;
#pragma unroll
      for (int j = 0; j < NLD; ++j) dst[j] = ld_poll_v4(src + (size_t)(2 * j) * Rp);     // strong load: see ld_poll_v4
    }
  };
  auto valid = [&](const float4 (&x)[NLD], int i, int p, uint32_t tag4) {
    const int sc = (int)(sched_at(sch, i, n_tiles, p) & 15u);
    const bool two = (a.KSLICE - sc * 64 >= 64);
    bool good = true;
    if (two || q < 8) {
#pragma unroll
      for (int j = 0; j < NLD; ++j) {
        const uint32_t bits = (__float_as_uint(x[j].x) & 1u) | ((__float_as_uint(x[j].y) & 1u) << 8) |
                              ((__float_as_uint(x[j].z) & 1u) << 16) | ((__float_as_uint(x[j].w) & 1u) << 24);
        good = good && (bits == tag4);
      }
    }
    return __all_sync(0xffffffffu, good);
  };
  for (int t = 0; t < T && okh; ++t)
    for (int k = 1; k < K && okh; ++k)
     for (int i = 0; i < n_tiles && okh; ++i, ++item) {
      const uint32_t tag4 = ll_tag(t, k - 1, K) * 0x01010101u;
      const int slot = (k - 1) & 1;
      bool have = false;                           // v holds (possibly stale) loads of the current position
      for (int p = 0; p < NSC && okh; ++p, hr.next(a.HST)) {
        if ((p & 1) != ci) continue;
        const int sc = (int)(sched_at(sch, i, n_tiles, p) & 15u);
        const bool two = (a.KSLICE - sc * 64 >= 64);
        const long long _t0 = clock64();
        unsigned spins = 0;
        if (!have) {
          const float* sp = a.hb_hi + ((size_t)slot * a.Bp + i * NB) * Rp + s * a.KSLICE + sc * 64 + 4 * q;
          for (;;) {                                  // sentinel: row 0 of the sub-chunk
            bool good = true;
            if (r0 == 0 && (two || q < 8)) {
              const float4 x = ld_poll_v4(sp);
              const uint32_t bits = (__float_as_uint(x.x) & 1u) | ((__float_as_uint(x.y) & 1u) << 8) |
                                    ((__float_as_uint(x.z) & 1u) << 16) | ((__float_as_uint(x.w) & 1u) << 24);
              good = bits == tag4;
            }
            if (__all_sync(0xffffffffu, good)) break;
            if ((++spins & 0x3F) == 0 && (clock64() - _t0 > RT_WATCHDOG || *err)) { okh = false; break; }
          }
          if (okh) issue(v, slot, i, p);
        }
        while (okh && !valid(v, i, p, tag4)) {
          if ((++spins & 0x3F) == 0 && (clock64() - _t0 > RT_WATCHDOG || *err)) { okh = false; break; }
          issue(v, slot, i, p);
        }
        okh = __all_sync(0xffffffffu, okh);
        if (!okh) {
          if (lane == 0) { atomicCAS(a.dev_error + 2, 0, (1 << 30) | (t << 16) | (k << 8) | (p << 4) | (ci << 1) | (have ? 1 : 0)); rt_fail(a.dev_error, 203); }
          break;
        }
        if (dbg_on) DBG_ADD(1, clock64() - _t0);
        if constexpr (PREFETCH) { if (p + 2 < NSC) issue(w, slot, i, p + 2); }     // in flight while this position is stored
        const int hs = hr.idx;
        if (lane == 0) {
          okh = mbar_wait(&bars->h_empty[hs], hr.ph ^ 1u, err, RT_WATCHDOG);
          if (!okh) rt_fail(a.dev_error, 202);
        }
        okh = __all_sync(0xffffffffu, okh);
        if (!okh) break;
        if (two || q < 8) {
          const uint32_t st0 = smem_u32(smem + a.off_h + hs * a.h_stage_bytes) + (uint32_t)(q >> 3) * HB;   // slab of this float4
#pragma unroll
          for (int j = 0; j < NLD; ++j) {
            const int r = r0 + 2 * j;
            const uint32_t ad = st0 + (uint32_t)r * 128u + (uint32_t)(((q & 7) ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "f"(v[j].x), "f"(v[j].y), "f"(v[j].z), "f"(v[j].w) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ad + 2 * HB), "f"(tf32_lo(v[j].x)), "f"(tf32_lo(v[j].y)),
                         "f"(tf32_lo(v[j].z)), "f"(tf32_lo(v[j].w)) : "memory");
          }
        }
        fence_proxy_async_smem();                // generic-proxy STS -> async-proxy reads of the MMA
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->h_full[hs]);
          RT_TRACE(1, 10 + p, item);
        }
        have = false;
        if constexpr (PREFETCH) {
          if (p + 2 < NSC) {
#pragma unroll
            for (int j = 0; j < NLD; ++j) v[j] = w[j];
            have = true;
          }
        }
      }
    }
}

template <int NB, bool BWD, int CB, bool SYM, bool SPLIT = false>
__global__ void __launch_bounds__(RT_THREADS, 1)
k_recurrent_tc(const __grid_constant__ CUtensorMap tmH_hi, const __grid_constant__ CUtensorMap tmH_lo,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW64,
               const __grid_constant__ RecSched sch_in, const RecArgs a_in) {
  // Batch groups: utterances are independent, so grid.z groups of (KS x MT) CTAs each run the whole chain on their own
  // contiguous range of batch tiles (disjoint SMs, own flags, no interaction).  Everything indexed by the utterance is
  // re-based once here; below, tile i / utterance b are group-local.
  RecArgs a = a_in;
  // pipelined forward: a CTA that executes is resident - the host holds the rest of the projection GEMM back until all
  // CTAs have counted themselves in (before any early return: the stream waits for the full count)
  if constexpr (NB <= 32) { if (a_in.started && threadIdx.x == 0) atomicAdd(a_in.started, 1u); }   // (latency plans only, see PIPE)
  // (the last group may hold fewer tiles: its schedule classes differ)
  SchedTab sch = (gridDim.z > 1 && blockIdx.z == gridDim.z - 1) ? sch_in.e_last : sch_in.e;
  const int tile0 = blockIdx.z * a_in.n_tiles;
  const int boff = tile0 * NB;
  if (gridDim.z > 1) {
    const size_t bo = (size_t)boff;
    a.n_tiles = min(a_in.n_tiles, a_in.n_tiles_total - tile0);
    a.B = min(a_in.B - boff, a.n_tiles * NB);
    a.mvalid += bo * a.T; a.flags += (size_t)tile0 * a.MT;
    a.hb_hi += bo * a.Rp; a.hb_lo += bo * a.Rp;
    if (!BWD) {
      a.XW += ((NB <= 32 && a.xw_tmajor) ? bo : bo * a.T) * a.K * a.Rp; a.state += bo * a.Rp; a.psum += bo;
      a.Hp_hi += bo * a.T * a.Rp; a.Hp_lo += bo * a.T * a.Rp;
      if (a.H_user) a.H_user += bo * a.T * a.R;
    } else {
      a.dH += bo * a.T * a.Rp; a.deltaT_hi += bo; a.deltaT_lo += bo; a.G += bo * a.Rp; a.psum2 += bo;
    }
    if (a.actT_hi) { a.actT_hi += bo; a.actT_lo += bo; }
  }
  // No static shared memory in this kernel: the dynamic window starts 1024-aligned (checked below).  Keeping `smem`
  // a plain __shared__ array (no integer round-trip) lets the compiler emit LDS/STS instead of generic LD/ST.
  extern __shared__ __align__(1024) uint8_t smem[];
  RecBars* bars = reinterpret_cast<RecBars*>(smem + a.off_bar);
  float* leak_s = reinterpret_cast<float*>(smem + a.off_leak);       // n_tiles x NB : sum_j state[b][j] of this frame
  float* out_s = reinterpret_cast<float*>(smem + a.off_out);         // NB x (RO+1) staging of the new state rows
  volatile int* err = a.dev_error;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) {               // uniform across the grid: everybody leaves
    if (threadIdx.x == 0) rt_fail(a.dev_error, 299);
    return;
  }
  const bool dbg_on = (a.dbg != nullptr) && blockIdx.x == 0 && blockIdx.y == a.dbg_m && blockIdx.z == 0;
  // reporting lanes: lane 0 of the single-warp roles, first thread of the four-warp roles
  const bool dbg_w = dbg_on && lane == 0 && (warp <= 4 || warp == 8 || warp == 12);
  const int dbg_wi = warp;
  const int s = blockIdx.x;            // K-split == rank in cluster
  const int m = blockIdx.y;            // M-tile
  const int K = a.K, T = a.T, Rp = a.Rp, n_tiles = a.n_tiles, NSC = a.NSC;
  constexpr uint32_t TMEM_COLS = 512;    // [128 b, 128 b + 64) W hi | [+64, +128) W lo of buffer b | [384, 384 + AST*NB) accumulators
  constexpr uint32_t ACC_COL0 = RT_WB * 128;
  constexpr uint32_t HB = NB * 128;      // one 32-atom x NB tile (hi or lo) of the hidden state

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmH_hi); tma_prefetch_desc(&tmH_lo); tma_prefetch_desc(&tmW);
    if (SYM) tma_prefetch_desc(&tmW64);
    for (int i = 0; i < RT_WB; ++i) { mbar_init(&bars->wb_full[i], 128); mbar_init(&bars->wb_empty[i], 1); }
    for (int i = 0; i < RT_MAXW; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_free[i], 4); }
    for (int i = 0; i < RT_MAXH; ++i) { mbar_init(&bars->h_full[i], 1); mbar_init(&bars->h_empty[i], 1); }
    for (int i = 0; i < RT_AST; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], 4); }
    for (int i = 0; i < 4; ++i) { mbar_init(&bars->red_full[i], 1); mbar_init(&bars->red_free[i], (SPLIT ? 8 : 4) * a.KS); }
    for (int i = 0; i < RT_PST; ++i) { mbar_init(&bars->pub_full[i], 4); mbar_init(&bars->pub_empty[i], 1); }
    bars->abort = 0;
    if (a.dbg) {
      for (int z = 0; z < 16 * 8; ++z) (&bars->dbg_acc[0][0])[z] = 0;
      for (int z = 0; z < 8; ++z) bars->dbg_trc[z] = 0;
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(&bars->tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // every CTA's barriers exist before any remote arrive / bulk copy
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const int n_mma_steps = T * (K - 1);
  if (dbg_w) bars->dbg_acc[dbg_wi][7] = clock64();

  if (warp == 0) {
    // ================= weight producer: TMA S_k^T[m*128 .. +128][32 K-columns] pieces into the smem ring ===============
    // The weights do not depend on the recurrence, so the ring runs ahead of the loaders that move them into TMEM
    // (a whole step ahead when one batch tile is in flight and the K-slice fits the ring).
    if (lane == 0) {
      const uint64_t pol_w = l2_policy_evict_last();
      RtRing wr;
      bool okw = true;
      for (int ms = 0; ms < n_mma_steps && okw; ++ms) {
        const int k = BWD ? (K - 1 - ms % (K - 1)) : (ms % (K - 1) + 1);
        for (int i = 0; i < n_tiles && okw; ++i)
          for (int p = 0; p < NSC && okw; ++p) {
            const uint32_t e = sched_at(sch, i, n_tiles, p);
            if (!(e & 64u)) continue;                                   // resident: nothing to load
            const int col0 = s * a.KSLICE + (int)(e & 15u) * 64;
            const int npc = (a.KSLICE - (int)(e & 15u) * 64 >= 64) ? 2 : 1;
            // Symmetric S_k: a sub-chunk whose K-columns lie in an M-tile below this CTA's own is fetched as the mirrored
            // block (rows = its K-columns, columns = this M-tile) in four 64 x 32 pieces, two per ring slot, and read
            // transposed by the loaders: only the blocks on and above the diagonal of a layer are ever touched (36 of 64
            // at R = 1000: 54 of 96 MB per frame), which is what stays resident in L2 across a frame.
            const bool mir = SYM && (col0 >> 7) < m;
            for (int pc = 0; pc < npc; ++pc, wr.next(a.WST)) {
              const int ws = wr.idx;
              RT_TIMED(0, okw = mbar_wait(&bars->w_free[ws], wr.ph ^ 1u, err, RT_WATCHDOG));
              if (!okw) { rt_fail(a.dev_error, 214); break; }
              mbar_expect_tx(&bars->w_full[ws], 16384u);
              if (mir) {
                tma_load_2d_hint(smem + a.off_w + ws * 16384, &tmW64, &bars->w_full[ws], m * 128 + (2 * pc) * 32, (k - 1) * Rp + col0, pol_w);
                tma_load_2d_hint(smem + a.off_w + ws * 16384 + 8192, &tmW64, &bars->w_full[ws], m * 128 + (2 * pc + 1) * 32, (k - 1) * Rp + col0, pol_w);
              } else {
                tma_load_2d_hint(smem + a.off_w + ws * 16384, &tmW, &bars->w_full[ws], col0 + pc * 32, (k - 1) * Rp + m * 128, pol_w);
              }
            }
          }
      }
    }
  } else if (warp == 1) {
    // ================= hidden-state loader: acquire the producers' flags, then TMA the sub-chunks of tile i ============
    // Flags are acquired lazily, right before the first sub-chunk that needs a producer M-tile (the M-tiles of a step
    // finish up to ~2k cycles apart): the lanes poll the (at most two) M-tiles a sub-chunk overlaps in parallel - a
    // satisfied poll still costs an L2 round trip -, then lane 0 issues one slab TMA per hi / lo.
    if (a.ll) {
      if constexpr (NB <= 32) ll_consumer<NB>(a, sch, bars, smem, 0, s, lane, dbg_on, dbg_w, dbg_wi);   // (the host never sets ll with NB = 64)
    } else {
      const int m_lo = (s * a.KSLICE) >> 7;
      RtRing hr;
      long long item = 0;
      bool okh = true;
      for (int t = 0; t < T && okh; ++t)
        for (int k = 1; k < K && okh; ++k) {
          const unsigned int target = (unsigned int)(a.pub_unit * a.KS) * (unsigned int)(t * K + k);   // step (t,k-1) published
          const int slot = (k - 1) & 1;
          for (int i = 0; i < n_tiles && okh; ++i, ++item) {
            unsigned int polled = 0;                 // producer M-tiles of this K-slice acquired for this item (warp-uniform)
            for (int p = 0; p < NSC && okh; ++p, hr.next(a.HST)) {
              const int sc = (int)(sched_at(sch, i, n_tiles, p) & 15u);
              const int c0 = s * a.KSLICE + sc * 64;
              const int wdt = (a.KSLICE - sc * 64 >= 64) ? 64 : 32;
              const int mt_a = c0 >> 7, mt_b = (c0 + wdt - 1) >> 7;
              const unsigned int need = (1u << (mt_a - m_lo)) | (1u << (mt_b - m_lo));
              if (need & ~polled) {
                bool mine_ok = true;
                const long long _t0 = dbg_on ? clock64() : 0;
                if (lane < 2) {
                  const int mt = lane == 0 ? mt_a : mt_b;
                  if (!((polled >> (mt - m_lo)) & 1u) && (lane == 0 || mt_b != mt_a))
                    mine_ok = poll_flag(a.flags + i * a.MT + mt, target, err);
                }
                okh = __all_sync(0xffffffffu, mine_ok);
                if (dbg_on) DBG_ADD(1, clock64() - _t0);
                if (!okh) { if (lane == 0) rt_fail(a.dev_error, 203); break; }
                polled |= need;
                if (lane == 0) fence_proxy_async_global();          // generic-proxy writes of the owners -> async-proxy (TMA) reads
                if (lane == 0) RT_TRACE(1, 1 + p, item);
              }
              if (lane == 0) {
                const int hs = hr.idx;
                RT_TIMED(0, okh = mbar_wait(&bars->h_empty[hs], hr.ph ^ 1u, err, RT_WATCHDOG));
                if (!okh) rt_fail(a.dev_error, 202);
                else {
                  uint8_t* dst = smem + a.off_h + hs * a.h_stage_bytes;
                  const int slab = c0 >> 5;
                  const int c1 = slot * a.Bp + boff + i * NB;          // the tensor map covers all groups
                  // stage = [hi slab 0 | hi slab 1 | lo slab 0 | lo slab 1]; a 32-wide last sub-chunk still moves two
                  // slabs (the second one is not multiplied; beyond the matrix it is zero-filled)
                  if (a.h3d) {
                    mbar_expect_tx(&bars->h_full[hs], 4u * HB);
                    tma_load_3d(dst, &tmH_hi, &bars->h_full[hs], 0, c1, slab);
                    tma_load_3d(dst + 2 * HB, &tmH_lo, &bars->h_full[hs], 0, c1, slab);
                  } else {                              // 2-D maps: one 32-atom tile per instruction
                    const int nat = wdt >> 5;
                    mbar_expect_tx(&bars->h_full[hs], (uint32_t)(2 * nat) * HB);
                    for (int at = 0; at < nat; ++at) {
                      tma_load_2d(dst + at * HB, &tmH_hi, &bars->h_full[hs], (slab + at) * 32, c1);
                      tma_load_2d(dst + (2 + at) * HB, &tmH_lo, &bars->h_full[hs], (slab + at) * 32, c1);
                    }
                  }
                  RT_TRACE(1, 10 + p, item);
                }
              }
              okh = __all_sync(0xffffffffu, okh);
            }
          }
        }
    }
  } else if (warp == 2) {
    // ================= MMA issuer =================
    // The whole warp runs the loop convergently (descriptor arithmetic stays in the uniform datapath); one elected
    // lane issues the tcgen05 instructions.
    {
      const uint32_t idesc = umma_idesc_tf32(128, NB);
      long long it = 0;
      RtRing hr, ar;
      uint32_t f0 = 0, f1 = 0, f2 = 0;               // fills of the three TMEM weight buffers seen so far
      int rot = 0;
      bool okm = true;
      const bool leader = elect_one();
      for (int ms = 0; ms < n_mma_steps && okm; ++ms) {
        for (int i = 0; i < n_tiles && okm; ++i, ++it, ar.next(RT_AST)) {
          const int as = ar.idx;
          RT_TIMED(0, okm = mbar_wait(&bars->t_empty[as], ar.ph ^ 1u, err, RT_WATCHDOG));
          if (!okm) { rt_fail(a.dev_error, 204); break; }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + ACC_COL0 + as * NB;
          for (int p = 0; p < NSC; ++p, hr.next(a.HST)) {
            const uint32_t e = sched_at(sch, i, n_tiles, p);
            const int sc = (int)(e & 15u);
            int wb = (int)((e >> 4) & 3u) + rot; wb = wb >= RT_WB ? wb - RT_WB : wb;
            if (e & 64u) {                           // first use: the loaders fill the buffer for this position
              const uint32_t f = wb == 0 ? f0 : (wb == 1 ? f1 : f2);
              RT_TIMED(2, okm = mbar_wait(&bars->wb_full[wb], f & 1u, err, RT_WATCHDOG));
              if (!okm) { rt_fail(a.dev_error, 206); break; }
              tc_fence_after();
              if (wb == 0) ++f0; else if (wb == 1) ++f1; else ++f2;
            }
            const int hs = hr.idx;
            RT_TIMED(1, okm = mbar_wait(&bars->h_full[hs], hr.ph, err, RT_WATCHDOG));
            if (!okm) { rt_fail(a.dev_error, 205); break; }
            // (no tcgen05.fence after this wait: the barrier is completed by TMA bytes, not by tcgen05 work)
            if (dbg_on && lane == 0 && p == 0) bars->dbg_ts[0] = clock64();
            if (lane == 0) RT_TRACE(2, 1 + p, it);
            const uint32_t wbase = tmem_base + (uint32_t)wb * 128u;
            const uint64_t dh = umma_desc_k128(smem_u32(smem + a.off_h + hs * a.h_stage_bytes));
            const bool two = (a.KSLICE - sc * 64 >= 64);
            // 12 MMAs per 32-atom slab, all descriptors = base + compile-time constants (34 cycles per N=64 MMA when the
            // queue never runs dry; address arithmetic per MMA cost 85)
#pragma unroll
            for (int at = 0; at < 2; ++at) {
              if (at == 1 && !two) break;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {       // the whole warp stays convergent, the issue itself is predicated
                const uint32_t w_hi = wbase + (at * 4 + kk) * 8, w_lo = wbase + 64 + (at * 4 + kk) * 8;
                const uint64_t h_hi = dh + (uint64_t)((at * HB + kk * 32) >> 4);
                const uint64_t h_lo = dh + (uint64_t)(((2 + at) * HB + kk * 32) >> 4);
                if (leader) umma_tf32_ts(d_tmem, w_lo, h_hi, idesc, (p | at | kk) != 0);
                if (leader) umma_tf32_ts(d_tmem, w_hi, h_lo, idesc, true);
                if (leader) umma_tf32_ts(d_tmem, w_hi, h_hi, idesc, true);
              }
            }
            if (leader) {
              tc_commit(&bars->h_empty[hs]);
              if (e & 128u) tc_commit(&bars->wb_empty[wb]);           // last use of this sub-chunk in the step
            }
            __syncwarp();
          }
          if (!okm) break;
          if (leader) tc_commit(&bars->t_full[as]);
          __syncwarp();
          if (lane == 0) RT_TRACE(2, 20, it);
          if (dbg_on && lane == 0) {   // acc3: first sub-chunk ready -> last MMA issued (the product phase)
            const long long _n = clock64(); DBG_ADD(3, _n - bars->dbg_ts[0]); bars->dbg_ts[1] = _n;
          }
        }
        rot += a.rot; rot = rot >= RT_WB ? rot - RT_WB : rot;
      }
    }
  } else if (warp >= 4 && warp < 8 && !SPLIT) {
    // ================= pushers: TMEM accumulator -> staging smem -> bulk DSMEM copy into every owner's slot =========
    // Thread rho holds accumulator row rho.  Rows are staged row-major with the 16-byte chunks of a row XOR-swizzled
    // by (row & 7) (conflict-free STS.128 here and LDS.128 in the owner); rows [o*RO, (o+1)*RO) are one contiguous
    // block that a single cp.async.bulk moves into owner o's reduction slot (block index = this CTA's rank).
    const int q = warp - 4;
    const int rho = q * 32 + lane;                   // accumulator row (TMEM lane) within the M-tile
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stage0 = smem_u32(smem + a.off_push);
    const uint32_t blk_bytes = (uint32_t)(a.RO * NB * 4);
    constexpr int CHUNKS = NB / 4;                   // 16-byte chunks per row
    constexpr int SWZ = (CHUNKS >= 8) ? 7 : CHUNKS - 1;
    int it = 0;
    RtRing ar, rr;
    bool dead = false;                               // a wait failed: run through (named barriers!) without touching mbarriers
    for (int ms = 0; ms < n_mma_steps; ++ms) {      // on a watchdog error the loop keeps running (waits return at once)
      for (int i = 0; i < n_tiles; ++i, ++it, ar.next(RT_AST), rr.next(a.RST)) {     // so that the named barrier below always sees all 128 threads
        const int as = ar.idx, rs = rr.idx;
        bool okp;
        RT_TIMED(0, okp = mbar_wait(&bars->t_full[as], ar.ph, err, RT_WATCHDOG));
        if (!okp) { rt_fail(a.dev_error, 207); dead = true; }
        tc_fence_after();
        if (threadIdx.x == 128) RT_TRACE(4, 1, it);
        if (dbg_on) DBG_ADD(2, clock64() - *reinterpret_cast<volatile long long*>(&bars->dbg_ts[1]));   // issue -> completion
        float v[NB];
#pragma unroll
        for (int c = 0; c < NB; c += 16) tmem_ld16(trow + ACC_COL0 + as * NB + c, v + c);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && !dead) mbar_arrive(&bars->t_empty[as]);
        // slot rs (staging here, reduction slot in the owners) is free once every owner has consumed its previous use
        RT_TIMED(1, okp = mbar_wait_cluster(&bars->red_free[rs], rr.ph ^ 1u, err, RT_WATCHDOG));
        if (!okp) { rt_fail(a.dev_error, 208); dead = true; }
        if (threadIdx.x == 128) RT_TRACE(4, 2, it);
        const uint32_t srow = stage0 + rs * a.red_slot_bytes + rho * (NB * 4);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const uint32_t addr = srow + (uint32_t)(((c & ~SWZ) | ((c ^ rho) & SWZ)) * 16);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                       "f"(v[4 * c + 2]), "f"(v[4 * c + 3]) : "memory");
        }
        fence_proxy_async_smem();                     // generic-proxy STS -> async-proxy bulk copy reads
        // owner o's block = rows [o*RO, (o+1)*RO): when a block lies inside one warp's 32 rows (RO <= 32) the warp
        // ships its own 32/RO blocks right away, otherwise the four warps meet first and warp 4 ships all of them
        int o_first, o_cnt;
        if (a.RO <= 32) { __syncwarp(); o_cnt = 32 / a.RO; o_first = q * o_cnt; }
        else { asm volatile("bar.sync 2, 128;" ::: "memory"); o_cnt = (warp == 4) ? a.KS : 0; o_first = 0; }
        if (lane < o_cnt && !dead) {                  // lane l copies this CTA's block for owner o_first + l
          const uint32_t o = (uint32_t)(o_first + lane);
          const uint32_t src = stage0 + rs * a.red_slot_bytes + o * blk_bytes;
          const uint32_t dst = mapa_u32(smem_u32(smem + a.off_red) + rs * a.red_slot_bytes + (uint32_t)s * blk_bytes, o);
          const uint32_t bar = mapa_u32(smem_u32(&bars->red_full[rs]), o);
          dsmem_bulk_copy(dst, src, blk_bytes, bar);
        }
        if (threadIdx.x == 128) RT_TRACE(4, 3, it);
      }
    }
  } else if (warp == 3) {
    // ================= publisher: one gpu-scope release per (step, tile) after the owners' stores =================
    // The owner threads only arrive on pub_full (release.cta); this thread acquires it, issues the single cumulative
    // gpu-scope fence and bumps flag[tile][m], so the owners never stall on a memory fence.
    if (a.ll) {
      if constexpr (NB <= 32) ll_consumer<NB>(a, sch, bars, smem, 1, s, lane, false, false, 3);
    } else if (lane == 0) {
      const long long n_items = (a.pub_unit == 1) ? (long long)T * K * n_tiles : 0;   // direct mode: the owners publish
      long long j = 0;
      while (j < n_items) {
        bool okb;
        RT_TIMED(0, okb = mbar_wait(&bars->pub_full[(int)(j % RT_PST)], (uint32_t)((j / RT_PST) & 1), err, RT_WATCHDOG));
        if (!okb) { rt_fail(a.dev_error, 211); break; }
        // batch every further item that is already complete behind ONE gpu-scope fence (throughput mode)
        long long jend = j + 1;
        while (jend < n_items && jend - j < RT_PST &&
               mbar_try_wait(&bars->pub_full[(int)(jend % RT_PST)], (uint32_t)((jend / RT_PST) & 1))) ++jend;
        // fence.acq_rel.gpu is cumulative over the owners' stores observed through the mbarriers (release/acquire.cta)
        RT_TIMED(1, asm volatile("fence.acq_rel.gpu;" ::: "memory"));
        for (long long q2 = j; q2 < jend; ++q2) {
          const int i = (int)(q2 % n_tiles);
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(a.flags + i * a.MT + m), "r"(1u) : "memory");
          mbar_arrive(&bars->pub_empty[(int)(q2 % RT_PST)]);
        }
        j = jend;
      }
    }
  } else if (!BWD && ((warp >= 8 && warp < 12) || (SPLIT && warp >= 4 && warp < 8))) {
    // ================= owners: reduce the KS partials of rows [o*RO, (o+1)*RO), epilogue, publish =================
    // Split epilogue (a.split: K-split 2 with 64 batch columns, 4096 outputs per item): the exchange of such an item is
    // a third of the K-split-4 one per flop, which leaves the owners as the bottleneck - so two groups of four warps own
    // half of the CTA's rows each.  Group 1 = warps 4-7, which first push the item (they are the only warps that can read
    // all 128 TMEM lanes besides the loaders), group 0 = warps 8-11.  Each group has its own leak / staging scratch, its
    // own named barrier and publishes as a CTA of its own (psum slot, 4 flag increments).
    constexpr int nparts_o = SPLIT ? 2 : 1;
    const int part_o = (SPLIT && warp < 8) ? 1 : 0;
    const int otid = threadIdx.x & 127;              // 0..127 inside the group
    const int RO = a.RO / nparts_o;                  // rows of this group
    const int rpart0 = part_o * RO;                  // their offset inside the CTA's block of the reduction slot
    const int row0 = m * 128 + s * a.RO + rpart0;    // first global output row of the group
    const int cta_lin = (m * a.KS + s) * nparts_o + part_o, n_cta = a.MT * a.KS * nparts_o;
    const int bar_id = part_o ? 3 : 1;
    float* leak_g = leak_s + part_o * (n_tiles * NB);
    float* out_g = out_s + part_o * (NB * (RO + 1));
    RtRing ar_push;                                  // accumulator stage of the next item to push (group 1)
    const size_t KRp = (size_t)K * Rp;
    // The RO x NB outputs of a tile are dealt to the 128 threads as ONE block of 4 rows x CB batch columns each
    // (CB = RO*NB/512, chosen by the host; fewer threads work when the tile is smaller).  Partials are read as
    // LDS.(32*CB) over the batch columns, results leave as float4 over the rows (K-major state).  Within a warp the
    // lanes cover 2 row-quads x 16 column groups: reads of a row are contiguous (conflict-free with the pusher's
    // swizzle) and the two row-quads fill whole 32-byte sectors of the stores.
    constexpr int CQ = NB / CB;                      // column groups per row
    constexpr int CQW = (CQ < 16) ? CQ : 16;
    constexpr int CHUNKS = NB / 4;
    constexpr int SWZ = (CHUNKS >= 8) ? 7 : CHUNKS - 1;
    const int n_blk = (RO / 4) * CQ;                 // a multiple of 32: whole warps are active or idle
    const bool mine = otid < n_blk;
    const int uu = mine ? otid : 0;
    const int my_cq = (uu % CQW) + CQW * ((uu / (2 * CQW)) % (CQ / CQW));
    const int my_rq = ((uu / CQW) % 2) + 2 * (uu / (2 * CQ));
    const int rowq = row0 + 4 * my_rq;
    // sum_j h0[j]: leak of frame 0 (state = h0 for every utterance), same fixed order in every CTA
    float h0sum = 0.f;
    for (int j = 0; j < a.R; ++j) h0sum += a.h0[j];
    // x~W_k + b_k of an item does not depend on the recurrence: fetched one item ahead, unconditionally from clamped
    // addresses, and only consumed an iteration later, so the in-order warp never waits on a load it just issued.
    const uint64_t pol_x = l2_policy_evict_first();
    auto fetch_xw = [&](int t, int k, int i, float4 (&xa)[CB]) {
      const int tc = t < T ? t : T - 1;
      if (mine) {
#pragma unroll
        for (int bi = 0; bi < CB; ++bi) {
          int b = i * NB + CB * my_cq + bi; b = b < a.B ? b : a.B - 1;
          // (the pipelined, time-major order exists for the latency plans only: the 64-column variants, whose owners
          //  bound the throughput regime, compile without it)
          const size_t xrow = (NB <= 32 && a.xw_tmajor) ? (size_t)tc * a.Btot + b : (size_t)b * T + tc;
          xa[bi] = ldg_hint4(a.XW + xrow * KRp + (size_t)k * Rp + rowq, pol_x);
        }
      }
    };
    float4 xa_next[CB];
#pragma unroll
    for (int bi = 0; bi < CB; ++bi) xa_next[bi] = make_float4(0.f, 0.f, 0.f, 0.f);
    fetch_xw(0, 0, 0, xa_next);
    int it = 0;
    RtRing rr;
    bool dead = false;                              // a wait of this thread failed: run through without touching mbarriers
    long long j = 0;
    for (int t = 0; t < T; ++t)
    for (int k = 0; k < K; ++k)
    for (int i = 0; i < n_tiles; ++i, ++j) {
      const bool last = (k == K - 1);
      const float dmo = (k == 0) ? a.u0_dmo : a.uk_dmo, off = (k == 0) ? a.u0_off : a.uk_off;
      float4 xw[CB];
      float acc[4][CB];                                // [row e][batch bi]
      long long _ts = dbg_on ? clock64() : 0;
#pragma unroll
      for (int bi = 0; bi < CB; ++bi) {
        xw[bi] = xa_next[bi];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e][bi] = 0.f;
      }
      {
        int i2 = i + 1, k2 = k, t2 = t;
        if (i2 == n_tiles) { i2 = 0; if (++k2 == K) { k2 = 0; ++t2; } }
        // pipelined forward: the projections of the frames >= xw_t0 are still being computed on other SMs when this
        // kernel starts; they are acquired once, before the first load that touches them
        if constexpr (NB <= 32) {
          if (a.xw_ready && t2 == a.xw_t0 && k2 == 0 && i2 == 0 && !poll_flag(a.xw_ready, 1u, err, 400000000LL)) rt_fail(a.dev_error, 215);
        }
        fetch_xw(t2, k2, i2, xa_next);
      }
      // Loads that do not depend on this item's product are issued BEFORE the wait and consumed after it:
      //  * k > 0: the identity part of S_k (see the weight loader) = this thread's own rows of g^{k-1}, written by it
      //    one layer earlier into the ping-pong buffer (raw fp32 in the "hi" half);
      //  * k == 0: the recurrent state (written by this thread at the last layer of the previous frame);
      //  * last layer: the Keras mask of the frame.
      float4 pre[CB];
      float mvp[CB];
#pragma unroll
      for (int bi = 0; bi < CB; ++bi) {
        pre[bi] = make_float4(0.f, 0.f, 0.f, 0.f);
        mvp[bi] = 0.f;
        if (mine) {
          const int bcol = i * NB + CB * my_cq + bi;
          const int bc = bcol < a.B ? bcol : a.B - 1;
          if (k > 0)
            pre[bi] = __ldcg(reinterpret_cast<const float4*>(a.hb_hi + ((size_t)((k - 1) & 1) * a.Bp + bcol) * Rp + rowq));
          else
            pre[bi] = (t == 0) ? __ldg(reinterpret_cast<const float4*>(a.h0 + rowq))
                               : __ldcg(reinterpret_cast<const float4*>(a.state + (size_t)bc * Rp + rowq));
          if (last) mvp[bi] = __ldcg(a.mvalid + (size_t)bc * T + t);   // L2: late frames' entries are written while the kernel runs
        }
      }
      if (dbg_on) { long long _n = clock64(); DBG_ADD(2, _n - _ts); _ts = _n; }
      if (k == 0) {
        // ---- frame start: leak[b] = sum_j state[b][j] from the published partial sums of the previous frame ----
        if (t > 0) {
          const unsigned int target = (unsigned int)(a.pub_unit * a.KS) * (unsigned int)(a.ll ? t : t * K);   // LL: flags only count frames
          if (otid < a.MT && !poll_flag(a.flags + i * a.MT + otid, target, err)) rt_fail(a.dev_error, 209);
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          const int b = otid % NB, part = otid / NB, nparts = 128 / NB;
          float sacc = 0.f;
          const float* ps = a.psum + (size_t)((t - 1) & 1) * 256 * a.Bp + i * NB + b;
#pragma unroll 16                                  // independent L2 loads in flight, summed in index order
          for (int c = part; c < n_cta; c += nparts) sacc += __ldcg(ps + (size_t)c * a.Bp);
          out_g[part * NB + b] = sacc;
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (otid < NB) {
            float tot = 0.f;
            for (int p = 0; p < nparts; ++p) tot += out_g[p * NB + otid];
            leak_g[i * NB + otid] = tot;
          }
        } else if (otid < NB) {
          leak_g[i * NB + otid] = h0sum;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      } else {
        // ---- wait for the KS partial tiles, sum them in rank order ----
        const int rs = rr.idx;
        // (after a failed wait the thread only runs through: re-arming a barrier whose phase never completed traps)
        if (SPLIT && part_o == 1) {                    // split epilogue: this group pushes the item first
          push_item<NB>(a, bars, smem, tmem_base, ACC_COL0, s, warp - 4, lane, ar_push.idx, ar_push.ph, rs, rr.ph, dead);
          ar_push.next(RT_AST);
        }
        if (otid == 0 && part_o == 0 && !dead) mbar_expect_tx(&bars->red_full[rs], (uint32_t)(128 * NB * 4));
        bool oko = false;
        if (!dead) RT_TIMED(0, oko = mbar_wait_cluster(&bars->red_full[rs], rr.ph, err, RT_WATCHDOG));
        if (!oko) { rt_fail(a.dev_error, 210); dead = true; }
        const uint32_t red = smem_u32(smem + a.off_red) + rs * a.red_slot_bytes;
        if (dbg_on) _ts = clock64();
        if (otid == 0) RT_TRACE(5, 1, it);
        // address of (row r, columns CB*cq..) inside a source block: row-major, 16-byte chunk index swizzled by
        // (row & 7); the pusher's row index rho = o*RO + r has the same low 3 bits as r because RO is a multiple of 8.
        const uint32_t src_stride = (uint32_t)(a.RO * NB * 4);
        if (mine) {
          uint32_t qaddr[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = rpart0 + 4 * my_rq + e, col = CB * my_cq, ch = col >> 2;
            qaddr[e] = red + (uint32_t)(r * (NB * 4) + ((ch & ~SWZ) | ((ch ^ r) & SWZ)) * 16 + (col & 3) * 4);
          }
#pragma unroll 4
          for (int src = 0; src < a.KS; ++src) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t ad = qaddr[e] + src * src_stride;
              if constexpr (CB == 4) {
                float x0, x1, x2, x3;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(ad));
                acc[e][0] += x0; acc[e][CB > 1 ? 1 : 0] += x1; acc[e][CB > 2 ? 2 : 0] += x2; acc[e][CB > 3 ? 3 : 0] += x3;
              } else if constexpr (CB == 2) {
                float x0, x1;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x0), "=f"(x1) : "r"(ad));
                acc[e][0] += x0; acc[e][CB > 1 ? 1 : 0] += x1;
              } else {
                float x0;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(ad));
                acc[e][0] += x0;
              }
            }
          }
        }
        __syncwarp();                                  // this warp has consumed the slot (values are in registers):
        if (lane < a.KS && !dead) mbar_arrive_remote_relaxed(&bars->red_free[rs], (uint32_t)lane);   // 4 warps x KS owners arrivals
        if (dbg_on) { long long _n = clock64(); DBG_ADD(3, _n - _ts); _ts = _n; }
        if (otid == 0) RT_TRACE(5, 2, it);
        ++it; rr.next(a.RST);
      }
      // ---- fused epilogue: relu(acc + x~W_k + b_k + leak terms), Keras mask carry on the last layer ----
      if (mine) {
        float gall[CB][4];                             // [batch bi][row e] kept for the transposed activation store
#pragma unroll
        for (int bi = 0; bi < CB; ++bi) {
          const int bl = CB * my_cq + bi, b = i * NB + bl;
          const float lkv = off * leak_g[i * NB + bl];
          const float pv[4] = {pre[bi].x, pre[bi].y, pre[bi].z, pre[bi].w};
          const float xv[4] = {xw[bi].x, xw[bi].y, xw[bi].z, xw[bi].w};
          float sv[4] = {0.f, 0.f, 0.f, 0.f};          // state entering the frame (only for U with d != o beyond layer 0)
          if (k > 0 && dmo != 0.f && b < a.B) {
            const float4 so = (t == 0) ? __ldg(reinterpret_cast<const float4*>(a.h0 + rowq))
                                       : __ldcg(reinterpret_cast<const float4*>(a.state + (size_t)b * Rp + rowq));
            sv[0] = so.x; sv[1] = so.y; sv[2] = so.z; sv[3] = so.w;
          }
          float g[4], stn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool valid = (b < a.B) && (rowq + e < a.R);
            // k == 0: pre = state, weighted by the diagonal excess of U_0;  k > 0: pre = g^{k-1} (identity of S_k)
            const float base = (k == 0) ? dmo * pv[e] : (pv[e] + dmo * sv[e]);
            g[e] = valid ? fmaxf(acc[e][bi] + xv[e] + lkv + base, 0.f) : 0.f;
            gall[bi][e] = g[e];
          }
          if (!last) {
            const size_t o2 = ((size_t)(k & 1) * a.Bp + b) * Rp + rowq;
            if (a.ll) {   // only the exchanged copy carries the tag (a tagged zero is a denormal: it must not reach the
                          // stored activations, whose sign the backward pass tests)
              const uint32_t tg = ll_tag(t, k, K);
              __stcg(reinterpret_cast<float4*>(a.hb_hi + o2), make_float4(ll_mark(g[0], tg), ll_mark(g[1], tg), ll_mark(g[2], tg), ll_mark(g[3], tg)));
            } else __stcg(reinterpret_cast<float4*>(a.hb_hi + o2), make_float4(g[0], g[1], g[2], g[3]));
            if (!a.ll) __stcg(reinterpret_cast<float4*>(a.hb_lo + o2), make_float4(tf32_lo(g[0]), tf32_lo(g[1]), tf32_lo(g[2]), tf32_lo(g[3])));
          } else if (b < a.B) {
            // Keras masked scan: out_t = m ? g : out_{t-1} (zeros before the first step); state = m ? g : state
            const size_t bt = (size_t)b * T + t;
            const bool mv = mvp[bi] != 0.f;
            float outv[4];
            if (mv) {
#pragma unroll
              for (int e = 0; e < 4; ++e) { outv[e] = g[e]; stn[e] = g[e]; }
            } else {
              float4 po = make_float4(0.f, 0.f, 0.f, 0.f);
              if (t > 0) po = __ldcg(reinterpret_cast<const float4*>(a.Hp_hi + (bt - 1) * Rp + rowq));
              const float4 ss = (t == 0) ? __ldg(reinterpret_cast<const float4*>(a.h0 + rowq))
                                         : __ldcg(reinterpret_cast<const float4*>(a.state + (size_t)b * Rp + rowq));
              outv[0] = po.x; outv[1] = po.y; outv[2] = po.z; outv[3] = po.w;
              stn[0] = ss.x; stn[1] = ss.y; stn[2] = ss.z; stn[3] = ss.w;
            }
            __stcg(reinterpret_cast<float4*>(a.Hp_hi + bt * Rp + rowq), make_float4(outv[0], outv[1], outv[2], outv[3]));
            __stcg(reinterpret_cast<float4*>(a.Hp_lo + bt * Rp + rowq),
                   make_float4(tf32_lo(outv[0]), tf32_lo(outv[1]), tf32_lo(outv[2]), tf32_lo(outv[3])));
            if (a.H_user) {
#pragma unroll
              for (int e = 0; e < 4; ++e) if (rowq + e < a.R) a.H_user[bt * a.R + rowq + e] = outv[e];
            }
            __stcg(reinterpret_cast<float4*>(a.state + (size_t)b * Rp + rowq), make_float4(stn[0], stn[1], stn[2], stn[3]));
          }
          if (last) {
#pragma unroll
            for (int e = 0; e < 4; ++e) out_g[bl * (RO + 1) + 4 * my_rq + e] = stn[e];
          }
        }
        if (a.actT_hi) {                               // backward needs every layer's post-relu output
          const size_t TB = (size_t)T * a.Bp;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const size_t o3 = ((size_t)k * Rp + rowq + e) * TB + (size_t)t * a.Bp + i * NB + CB * my_cq;
            if constexpr (CB == 4) {
              __stcg(reinterpret_cast<float4*>(a.actT_hi + o3), make_float4(gall[0][e], gall[CB > 1 ? 1 : 0][e], gall[CB > 2 ? 2 : 0][e], gall[CB > 3 ? 3 : 0][e]));
              __stcg(reinterpret_cast<float4*>(a.actT_lo + o3), make_float4(tf32_lo(gall[0][e]), tf32_lo(gall[CB > 1 ? 1 : 0][e]),
                                                                           tf32_lo(gall[CB > 2 ? 2 : 0][e]), tf32_lo(gall[CB > 3 ? 3 : 0][e])));
            } else if constexpr (CB == 2) {
              __stcg(reinterpret_cast<float2*>(a.actT_hi + o3), make_float2(gall[0][e], gall[CB > 1 ? 1 : 0][e]));
              __stcg(reinterpret_cast<float2*>(a.actT_lo + o3), make_float2(tf32_lo(gall[0][e]), tf32_lo(gall[CB > 1 ? 1 : 0][e])));
            } else {
              __stcg(a.actT_hi + o3, gall[0][e]);
              __stcg(a.actT_lo + o3, tf32_lo(gall[0][e]));
            }
          }
        }
      }
      if (last) {
        // partial row sums of the new state over this CTA's RO rows (rank-1 leak of the next frame)
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (otid < NB) {
          float ps = 0.f;
          for (int r = 0; r < RO; ++r) ps += out_g[otid * (RO + 1) + r];
          __stcg(a.psum + (size_t)(t & 1) * 256 * a.Bp + (size_t)cta_lin * a.Bp + i * NB + otid, ps);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // out_s may be rewritten by the next item
      }
      if (dbg_on) {   // epilogue time by item kind: frame start (k = 0, incl. the psum flag wait) | last layer | regular
        const long long _n = clock64(), _d = _n - _ts; _ts = _n;
        if (k == 0) DBG_ADD(4, _d); else if (last) DBG_ADD(6, _d); else DBG_ADD(5, _d);
      }
      if (a.pub_unit != 1) {
        // latency mode (one batch tile): every owner warp releases its own stores - ONE wait for the write acks on the
        // critical path instead of release.cta arrive + the publisher's gpu fence back to back
        __syncwarp();
        if (otid == 0 && k > 0) RT_TRACE(5, 3, it - 1);
        // LL exchange: the hidden rows validate themselves; only the frame end (state / partial row sums) is flagged
        if (lane == 0 && (!a.ll || last)) RT_TIMED(1, flag_add_release(a.flags + i * a.MT + m, 1u));
        if (otid == 0 && k > 0) RT_TRACE(5, 4, it - 1);
      } else {
        // ---- hand the item to the publisher (release.cta arrive; the publisher's fence makes it gpu-visible) ----
        const int ps_ = (int)(j % RT_PST);
        bool okq;
        RT_TIMED(1, okq = mbar_wait(&bars->pub_empty[ps_], (uint32_t)(((j / RT_PST) & 1) ^ 1), err, RT_WATCHDOG));
        if (!okq) rt_fail(a.dev_error, 212);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pub_full[ps_]);   // release.cta, cumulative over the warp's stores (after __syncwarp)
      }
    }
  }
  if (BWD && warp >= 8 && warp < 12) {
    // ================= owners, BACKWARD chain =================
    // Same item structure as the forward pass with u = position in the frame: u = 0 builds delta^{K-1}(t) from the head
    // gradient and the carried state gradient G (needs the row sums of the previous processed frame, exchanged through
    // psum2 like the forward leak); u >= 1 turns the reduced product delta^k . S_k into delta^{k-1} by masking with the
    // stored activation.  Every delta is also written transposed (time-major) for the weight-gradient GEMMs.
    // Work split as in the forward owners: one block of 4 rows x CB batch columns per thread, all loads that do not
    // depend on the item's product issued before the wait.
    const int otid = threadIdx.x - 256;
    const int RO = a.RO;
    const int row0 = m * 128 + s * RO;
    const int cta_lin = m * a.KS + s, n_cta = a.MT * a.KS;
    constexpr int CQ = NB / CB;
    constexpr int CQW = (CQ < 16) ? CQ : 16;
    constexpr int CHUNKS = NB / 4;
    constexpr int SWZ = (CHUNKS >= 8) ? 7 : CHUNKS - 1;
    const int RQ = RO / 4;
    const int n_blk = RQ * CQ;
    const bool mine = otid < n_blk;
    const int uu = mine ? otid : 0;
    const int my_cq = (uu % CQW) + CQW * ((uu / (2 * CQW)) % (CQ / CQW));
    const int my_rq = ((uu / CQW) % 2) + 2 * (uu / (2 * CQ));
    const int rowq = row0 + 4 * my_rq;
    const int col0b = CB * my_cq;                            // first batch column of the block inside the tile
    const size_t TB = (size_t)T * a.Bp;
    float* add_s = leak_s;                                   // n_tiles x NB : o0*R0 + ok*Rk of the previous processed frame
    float* rk_acc = leak_s + n_tiles * NB;                   // n_tiles x NB : running sum over layers >= 1 of this frame
    float* part_s = out_s;                                   // 2 x RQ x NB  : per-item row-sum partials (double buffered)
    auto fetch_act = [&](int fi, int u, int i, float (&av)[4][CB]) {
      const int fic = fi < T ? fi : T - 1;
      const int t = T - 1 - fic;
      const int la = (u == 0) ? K - 1 : K - u - 1;
      if (mine) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float* src = a.actT_hi + ((size_t)la * Rp + rowq + e) * TB + (size_t)t * a.Bp + i * NB + col0b;
          if constexpr (CB == 4) {
            const float4 f = __ldcg(reinterpret_cast<const float4*>(src));
            av[e][0] = f.x; av[e][CB > 1 ? 1 : 0] = f.y; av[e][CB > 2 ? 2 : 0] = f.z; av[e][CB > 3 ? 3 : 0] = f.w;
          } else if constexpr (CB == 2) {
            const float2 f = __ldcg(reinterpret_cast<const float2*>(src));
            av[e][0] = f.x; av[e][CB > 1 ? 1 : 0] = f.y;
          } else {
            av[e][0] = __ldcg(src);
          }
        }
      }
    };
    float act_next[4][CB];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int bi = 0; bi < CB; ++bi) act_next[e][bi] = 0.f;
    fetch_act(0, 0, 0, act_next);
    int it = 0;
    RtRing rr;
    bool dead = false;                              // a wait of this thread failed: run through without touching mbarriers
    long long j = 0;
    for (int fi = 0; fi < T; ++fi)
    for (int u = 0; u < K; ++u)
    for (int i = 0; i < n_tiles; ++i, ++j) {
      const int t = T - 1 - fi;
      const int la = (u == 0) ? K - 1 : K - u - 1;           // layer of the delta this item produces
      float av[4][CB];
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int bi = 0; bi < CB; ++bi) av[e][bi] = act_next[e][bi];
      {
        int i2 = i + 1, u2 = u, f2 = fi;
        if (i2 == n_tiles) { i2 = 0; if (++u2 == K) { u2 = 0; ++f2; } }
        fetch_act(f2, u2, i2, act_next);
      }
      // ---- loads that do not depend on this item's product (this thread wrote them itself, or they are inputs) ----
      // Vector alph: S_k^T = I - diag(1/alph_k) G (G symmetric) is not symmetric, but
      //   sum_j delta_j S_k[i][j] = alph_k,i * sum_j (delta_j / alph_k,j) S_k^T[i][j],
      // so the stored matrix still serves: the operand is written pre-scaled by 1/alph and the product is scaled back.
      float4 a_post = make_float4(1.f, 1.f, 1.f, 1.f), a_pre = a_post;   // alph of the product's layer / 1/alph of the next one
      if (a.alph_vec && mine) {
        if (u > 0) a_post = __ldg(reinterpret_cast<const float4*>(a.alph_vec + (size_t)(K - u) * Rp + rowq));
        if (la > 0) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(a.alph_vec + (size_t)la * Rp + rowq));
          a_pre = make_float4(1.f / t4.x, 1.f / t4.y, 1.f / t4.z, 1.f / t4.w);
        }
      }
      float4 pre[CB];                                         // u > 0: own rows of delta^k (identity of S_k); u == 0: G
      float4 dh4[CB];                                         // u == 0: head gradient
      float mvt[CB], mvn[CB];                                 // mask of frame t / t+1
#pragma unroll
      for (int bi = 0; bi < CB; ++bi) {
        pre[bi] = make_float4(0.f, 0.f, 0.f, 0.f); dh4[bi] = pre[bi]; mvt[bi] = 0.f; mvn[bi] = 0.f;
        const int b = i * NB + col0b + bi;
        if (mine) {
          if (u > 0) {
            pre[bi] = __ldcg(reinterpret_cast<const float4*>(a.hb_hi + ((size_t)((u - 1) & 1) * a.Bp + b) * Rp + rowq));
          } else if (b < a.B) {
            const size_t bt = (size_t)b * T + t;
            dh4[bi] = __ldg(reinterpret_cast<const float4*>(a.dH + bt * Rp + rowq));
            if (fi > 0) {
              pre[bi] = __ldcg(reinterpret_cast<const float4*>(a.G + (size_t)b * Rp + rowq));
              mvn[bi] = __ldg(a.mvalid + bt + 1);
            }
          }
          if ((u == 0 || la == 0) && b < a.B) mvt[bi] = __ldg(a.mvalid + (size_t)b * T + t);
        }
      }
      float d[4][CB];                                         // [row e][batch bi]
      if (u == 0) {
        if (fi > 0) {
          const unsigned int target = (unsigned int)(a.pub_unit * a.KS) * (unsigned int)(a.ll ? fi : fi * K);
          if (otid < a.MT && !poll_flag(a.flags + i * a.MT + otid, target, err)) rt_fail(a.dev_error, 219);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          // every CTA's stores of the frames processed so far (fi of them) have been released and acquired here: tell the
          // host-side streams that the weight-gradient GEMMs over those frames may start (cumulative release)
          if (a.progress && otid == 0 && s == 0 && m == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.progress), "r"((unsigned int)fi) : "memory");
          const int b = otid % NB, part = otid / NB, nparts = 128 / NB;
          float s0 = 0.f, s1 = 0.f;
          const float* ps = a.psum2 + (size_t)((fi - 1) & 1) * 256 * 2 * a.Bp + i * NB + b;
#pragma unroll 8
          for (int c = part; c < n_cta; c += nparts) { s0 += __ldcg(ps + (size_t)c * 2 * a.Bp); s1 += __ldcg(ps + (size_t)c * 2 * a.Bp + a.Bp); }
          part_s[part * NB + b] = a.o0_b * s0 + a.ok_b * s1;
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (otid < NB) {
            float tot = 0.f;
            for (int p = 0; p < nparts; ++p) tot += part_s[p * NB + otid];
            add_s[i * NB + otid] = tot;
          }
        }
        if (otid < NB) rk_acc[i * NB + otid] = 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int bi = 0; bi < CB; ++bi) {
          const int bl = col0b + bi, b = i * NB + bl;
          float4 g4 = pre[bi];
          float dg[4] = {0.f, 0.f, 0.f, 0.f};
          if (mine && b < a.B) {
            if (fi > 0 && mvn[bi] != 0.f) {                   // frame t+1 was valid: its state gradient is complete now
              const float ad = add_s[i * NB + bl];
              g4.x = (rowq + 0 < a.R) ? g4.x + ad : 0.f; g4.y = (rowq + 1 < a.R) ? g4.y + ad : 0.f;
              g4.z = (rowq + 2 < a.R) ? g4.z + ad : 0.f; g4.w = (rowq + 3 < a.R) ? g4.w + ad : 0.f;
              __stcg(reinterpret_cast<float4*>(a.G + (size_t)b * Rp + rowq), g4);
            }
            if (mvt[bi] != 0.f) {
              dg[0] = dh4[bi].x + g4.x; dg[1] = dh4[bi].y + g4.y; dg[2] = dh4[bi].z + g4.z; dg[3] = dh4[bi].w + g4.w;
            }
          }
          d[0][bi] = dg[0]; d[1][bi] = dg[1]; d[2][bi] = dg[2]; d[3][bi] = dg[3];
        }
      } else {
        const int rs = rr.idx;
        if (otid == 0 && !dead) mbar_expect_tx(&bars->red_full[rs], (uint32_t)(128 * NB * 4));
        if (dead || !mbar_wait_cluster(&bars->red_full[rs], rr.ph, err, RT_WATCHDOG)) { rt_fail(a.dev_error, 220); dead = true; }
        const uint32_t red = smem_u32(smem + a.off_red) + rs * a.red_slot_bytes;
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int bi = 0; bi < CB; ++bi) d[e][bi] = 0.f;
        const uint32_t src_stride = (uint32_t)(RO * NB * 4);
        if (mine) {
          uint32_t qaddr[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = 4 * my_rq + e, ch = col0b >> 2;
            qaddr[e] = red + (uint32_t)(r * (NB * 4) + ((ch & ~SWZ) | ((ch ^ r) & SWZ)) * 16 + (col0b & 3) * 4);
          }
#pragma unroll 4
          for (int src = 0; src < a.KS; ++src) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t ad = qaddr[e] + src * src_stride;
              if constexpr (CB == 4) {
                float x0, x1, x2, x3;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(ad));
                d[e][0] += x0; d[e][CB > 1 ? 1 : 0] += x1; d[e][CB > 2 ? 2 : 0] += x2; d[e][CB > 3 ? 3 : 0] += x3;
              } else if constexpr (CB == 2) {
                float x0, x1;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x0), "=f"(x1) : "r"(ad));
                d[e][0] += x0; d[e][CB > 1 ? 1 : 0] += x1;
              } else {
                float x0;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(ad));
                d[e][0] += x0;
              }
            }
          }
        }
        __syncwarp();
        if (lane < a.KS && !dead) mbar_arrive_remote_relaxed(&bars->red_free[rs], (uint32_t)lane);
        ++it; rr.next(a.RST);
#pragma unroll
        for (int bi = 0; bi < CB; ++bi) {   // identity part of S_k: this owner's rows of the operand; scale back (vector alph)
          d[0][bi] = (d[0][bi] + pre[bi].x) * a_post.x; d[1][bi] = (d[1][bi] + pre[bi].y) * a_post.y;
          d[2][bi] = (d[2][bi] + pre[bi].z) * a_post.z; d[3][bi] = (d[3][bi] + pre[bi].w) * a_post.w;
        }
      }
      // ---- mask with the stored activation, store both layouts, row-sum partials ----
      if (mine) {
        float psb[CB];
#pragma unroll
        for (int bi = 0; bi < CB; ++bi) psb[bi] = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int bi = 0; bi < CB; ++bi) {
            const bool keep = (av[e][bi] > 0.f) && (rowq + e < a.R) && (i * NB + col0b + bi < a.B);
            d[e][bi] = keep ? d[e][bi] : 0.f;
            psb[bi] += d[e][bi];
          }
#pragma unroll
        for (int e = 0; e < 4; ++e) {                          // transposed (time-major) copy for the weight gradients
          const size_t o3 = ((size_t)la * Rp + rowq + e) * TB + (size_t)t * a.Bp + i * NB + col0b;
          if constexpr (CB == 4) {
            __stcg(reinterpret_cast<float4*>(a.deltaT_hi + o3), make_float4(d[e][0], d[e][CB > 1 ? 1 : 0], d[e][CB > 2 ? 2 : 0], d[e][CB > 3 ? 3 : 0]));
            __stcg(reinterpret_cast<float4*>(a.deltaT_lo + o3), make_float4(tf32_lo(d[e][0]), tf32_lo(d[e][CB > 1 ? 1 : 0]),
                                                                             tf32_lo(d[e][CB > 2 ? 2 : 0]), tf32_lo(d[e][CB > 3 ? 3 : 0])));
          } else if constexpr (CB == 2) {
            __stcg(reinterpret_cast<float2*>(a.deltaT_hi + o3), make_float2(d[e][0], d[e][CB > 1 ? 1 : 0]));
            __stcg(reinterpret_cast<float2*>(a.deltaT_lo + o3), make_float2(tf32_lo(d[e][0]), tf32_lo(d[e][CB > 1 ? 1 : 0])));
          } else {
            __stcg(a.deltaT_hi + o3, d[e][0]);
            __stcg(a.deltaT_lo + o3, tf32_lo(d[e][0]));
          }
        }
#pragma unroll
        for (int bi = 0; bi < CB; ++bi) {
          const int b = i * NB + col0b + bi;
          if (la > 0) {                                        // operand of the next product (pre-scaled for vector alph)
            const size_t o2 = ((size_t)(u & 1) * a.Bp + b) * Rp + rowq;
            float o0v = d[0][bi] * a_pre.x, o1v = d[1][bi] * a_pre.y, o2v = d[2][bi] * a_pre.z, o3v = d[3][bi] * a_pre.w;
            if (a.ll) { const uint32_t tg = ll_tag(fi, u, K); o0v = ll_mark(o0v, tg); o1v = ll_mark(o1v, tg); o2v = ll_mark(o2v, tg); o3v = ll_mark(o3v, tg); }
            __stcg(reinterpret_cast<float4*>(a.hb_hi + o2), make_float4(o0v, o1v, o2v, o3v));
            if (!a.ll) __stcg(reinterpret_cast<float4*>(a.hb_lo + o2), make_float4(tf32_lo(o0v), tf32_lo(o1v), tf32_lo(o2v), tf32_lo(o3v)));
          } else if (b < a.B && mvt[bi] != 0.f) {
            // delta^0: start of the new state gradient  G = (d0-o0) delta^0 (+ rank-1 terms added at the next frame start)
            __stcg(reinterpret_cast<float4*>(a.G + (size_t)b * Rp + rowq),
                   make_float4(a.d0mo_b * d[0][bi], a.d0mo_b * d[1][bi], a.d0mo_b * d[2][bi], a.d0mo_b * d[3][bi]));
          }
        }
        const int pb = (int)(j & 1);
#pragma unroll
        for (int bi = 0; bi < CB; ++bi) part_s[(pb * RQ + my_rq) * NB + col0b + bi] = psb[bi];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (otid < NB) {
        const int pb = (int)(j & 1);
        float sum = 0.f;
        for (int q2 = 0; q2 < RQ; ++q2) sum += part_s[(pb * RQ + q2) * NB + otid];
        if (la > 0) rk_acc[i * NB + otid] += sum;
        if (la == 0 || K == 1) {
          float* ps = a.psum2 + (size_t)(fi & 1) * 256 * 2 * a.Bp + (size_t)cta_lin * 2 * a.Bp + i * NB + otid;
          __stcg(ps, (la == 0) ? sum : 0.f);
          __stcg(ps + a.Bp, rk_acc[i * NB + otid]);
        }
      }
      if (a.pub_unit != 1) {
        __syncwarp();
        if (lane == 0 && (!a.ll || la == 0 || K == 1)) flag_add_release(a.flags + i * a.MT + m, 1u);
      } else {
        const int ps_ = (int)(j % RT_PST);
        if (!mbar_wait(&bars->pub_empty[ps_], (uint32_t)(((j / RT_PST) & 1) ^ 1), err, RT_WATCHDOG)) rt_fail(a.dev_error, 222);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pub_full[ps_]);
      }
    }
  }
  if (warp >= 12) {
    // ================= weight loaders: smem ring -> registers -> (hi | lo) in a TMEM weight buffer =================
    // Thread = row of the M-tile = TMEM lane.  A piece is 128 rows x 128 bytes in the TMA 128B-swizzle layout (16-byte
    // chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4)): eight consecutive rows read eight different chunks, so the
    // LDS.128 of a quarter-warp are conflict-free.  The loaders walk the same schedule as the MMA warp and fill a buffer
    // whenever a position is a first use.
    const int q = warp - 12;                          // TMEM lane quarter (warp % 4)
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    RtRing wr;
    uint32_t f0 = 0, f1 = 0, f2 = 0;
    int rot = 0;
    bool okl = true;
    for (int ms = 0; ms < n_mma_steps && okl; ++ms) {
      for (int i = 0; i < n_tiles && okl; ++i)
        for (int p = 0; p < NSC && okl; ++p) {
          const uint32_t e = sched_at(sch, i, n_tiles, p);
          if (!(e & 64u)) continue;
          const int sc = (int)(e & 15u);
          int wb = (int)((e >> 4) & 3u) + rot; wb = wb >= RT_WB ? wb - RT_WB : wb;
          const int col0 = s * a.KSLICE + sc * 64;
          const int npc = (a.KSLICE - sc * 64 >= 64) ? 2 : 1;
          const uint32_t tb = trow + (uint32_t)wb * 128u;
          if (SYM && (col0 >> 7) < m) {
            // mirrored sub-chunk: piece q (two per ring slot) = [64 K-columns (rows)] x [output rows 32q .. 32q+31 (128 bytes)].
            // Warp q owns exactly those TMEM lanes and reads its piece transposed: for a fixed K-column the 32 lanes read
            // one 128-byte row (conflict-free).  No diagonal in these blocks.  Every warp waits for and hands back both
            // slots so that the arrival counts of consecutive sub-chunks cannot mix.
            const int ws0 = wr.idx; const uint32_t ph0 = wr.ph; wr.next(a.WST);
            const int ws1 = wr.idx; const uint32_t ph1 = wr.ph; wr.next(a.WST);
            RT_TIMED(1, okl = mbar_wait(&bars->w_full[ws0], ph0, err, RT_WATCHDOG) && mbar_wait(&bars->w_full[ws1], ph1, err, RT_WATCHDOG));
            if (!okl) { rt_fail(a.dev_error, 215); break; }
            const uint32_t piece = smem_u32(smem + a.off_w + ((q >> 1) ? ws1 : ws0) * 16384 + (q & 1) * 8192);
            const uint32_t lcol = (uint32_t)(lane & 3) * 4u, lchunk = (uint32_t)(lane >> 2);
            {
              const uint32_t f = wb == 0 ? f0 : (wb == 1 ? f1 : f2);
              RT_TIMED(0, okl = mbar_wait(&bars->wb_empty[wb], (f & 1u) ^ 1u, err, RT_WATCHDOG));
              if (!okl) { rt_fail(a.dev_error, 213); break; }
              tc_fence_after();
            }
            for (int half = 0; half < 2; ++half) {
              float v[32], lo[32];
#pragma unroll
              for (int e2 = 0; e2 < 32; ++e2) {
                const uint32_t c = (uint32_t)(half * 32 + e2);
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[e2]) : "r"(piece + c * 128u + ((lchunk ^ (c & 7u)) << 4) + lcol));
              }
#pragma unroll
              for (int e2 = 0; e2 < 32; ++e2) lo[e2] = tf32_lo(v[e2]);
              tmem_st32(tb + half * 32, v);
              tmem_st32(tb + 64 + half * 32, lo);
            }
            fence_proxy_async_smem();                // all values have been consumed by the arithmetic above
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bars->w_free[ws0]); mbar_arrive(&bars->w_free[ws1]); }
          } else
          for (int pc = 0; pc < npc; ++pc, wr.next(a.WST)) {
            const int ws = wr.idx;
            RT_TIMED(1, okl = mbar_wait(&bars->w_full[ws], wr.ph, err, RT_WATCHDOG));
            if (!okl) { rt_fail(a.dev_error, 215); break; }
            const uint32_t rbase = smem_u32(smem + a.off_w + ws * 16384) + (uint32_t)row * 128u;
            float v[32];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4 * c4]), "=f"(v[4 * c4 + 1]), "=f"(v[4 * c4 + 2]),
                           "=f"(v[4 * c4 + 3]) : "r"(rbase + (uint32_t)((c4 ^ (row & 7)) << 4)));
            if (pc == 0) {
              const uint32_t f = wb == 0 ? f0 : (wb == 1 ? f1 : f2);
              RT_TIMED(0, okl = mbar_wait(&bars->wb_empty[wb], (f & 1u) ^ 1u, err, RT_WATCHDOG));
              if (!okl) { rt_fail(a.dev_error, 213); break; }
              tc_fence_after();
            }
            // The tensor core multiplies by S_k^T - I = -(G_k)^T only; the identity part of the product (the operand
            // itself) is added exactly, in fp32, by the row owner.  tcgen05 accumulates with round-toward-zero, and with
            // the identity inside the product that bias (relative to |h|) piles up over the K_layers x T chain (measured
            // 1.1e-4 on H at R=1000, K=25 against 5e-6 for this form).
            if (col0 + pc * 32 == m * 128 + q * 32) {     // warp-uniform: this piece holds the diagonal, in column `lane`
#pragma unroll
              for (int e2 = 0; e2 < 32; ++e2) v[e2] -= (e2 == lane) ? 1.0f : 0.0f;   // select, not a branch (a 32-way jump table otherwise)
            }
            float lo[32];
#pragma unroll
            for (int e2 = 0; e2 < 32; ++e2) lo[e2] = tf32_lo(v[e2]);
            // Hand the piece back only now: the arithmetic above depends on every loaded value, so the LDS have really
            // completed (an arrive issued right behind the loads let the refill overtake them: measured, non-repeatable
            // results), and the generic-proxy reads are fenced against the async-proxy (TMA) refill.
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->w_free[ws]);
            tmem_st32(tb + pc * 32, v);
            tmem_st32(tb + 64 + pc * 32, lo);
          }
          if (!okl) break;
          RT_TIMED(5, tc_wait_st());
          tc_fence_before();
          mbar_arrive(&bars->wb_full[wb]);
          if (threadIdx.x == 384) RT_TRACE(6, 1 + sc, (long long)ms * n_tiles + i);
          if (wb == 0) ++f0; else if (wb == 1) ++f1; else ++f2;
        }
      rot += a.rot; rot = rot >= RT_WB ? rot - RT_WB : rot;
    }
  }
  if (dbg_w) {
    long long* d = a.dbg + dbg_wi * 8;                    // slot per warp: [total, acc0..acc6]
    d[0] = clock64() - bars->dbg_acc[dbg_wi][7];
    for (int q2 = 0; q2 < 7; ++q2) d[1 + q2] = bars->dbg_acc[dbg_wi][q2];
  }
  // ---- teardown: nobody leaves while a peer may still touch its shared memory ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<TMEM_COLS>(tmem_base); }
}

// ---------------------------------------------------------------------------------------------------
// n_tiles = batch tiles per group, G = batch groups (grid.z), n_tiles_total = tiles of the whole batch
struct RecPlan { int NB, KS, MT, RO, NSC, KSLICE, n_tiles, n_tiles_total, G, WST, HST, RST, CB; size_t smem; RecArgs a; RecSched sch; bool ok; const char* why; };

// Per-step schedule (see RecSched).  NSC <= RT_WB: every sub-chunk of the step is resident, all tiles walk the K-slice
// upwards (results do not depend on the tile an utterance lands in) and the buffers rotate from step to step so that the
// next step's first sub-chunks are converted while this step still multiplies.  NSC > RT_WB: buffer = sub-chunk mod 3,
// tiles walk the K-slice in alternating direction and reuse whatever the previous tile left resident.
static void build_schedule(uint8_t (*sce)[RT_MAXSC], int NSC, int n_tiles, int* rot) {
  memset(sce, 0, 6 * RT_MAXSC);
  if (NSC <= RT_WB) {
    *rot = NSC % RT_WB;
    for (int cls = 0; cls < 6; ++cls)
      for (int p = 0; p < NSC; ++p)
        sce[cls][p] = (uint8_t)(p | (p << 4) | ((cls % 3 == 0) ? 64 : 0) | ((cls >= 3) ? 128 : 0));
    return;
  }
  *rot = 0;
  const int n = n_tiles < 6 ? n_tiles : 6 - ((n_tiles & 1) ? 1 : 0);   // the pattern has period 2: simulate a short step of the same parity
  struct Acc { int sc, slot, fresh, last; };
  Acc acc[6 * RT_MAXSC];
  int resident[RT_WB] = {-1, -1, -1};
  for (int i = 0; i < n; ++i)
    for (int p = 0; p < NSC; ++p) {
      Acc& x = acc[i * NSC + p];
      x.sc = (i & 1) ? NSC - 1 - p : p;
      x.slot = x.sc % RT_WB;
      x.fresh = resident[x.slot] != x.sc;
      resident[x.slot] = x.sc;
      x.last = 1;
    }
  for (int j = 0; j < n * NSC; ++j)
    for (int j2 = j + 1; j2 < n * NSC; ++j2)
      if (acc[j2].slot == acc[j].slot) { acc[j].last = acc[j2].sc != acc[j].sc; break; }
  for (int i = 0; i < n; ++i) {
    const int cls = (i == 0 ? 0 : ((i & 1) ? 1 : 2)) + (i == n - 1 ? 3 : 0);
    for (int p = 0; p < NSC; ++p) {
      const Acc& x = acc[i * NSC + p];
      sce[cls][p] = (uint8_t)(x.sc | (x.slot << 4) | (x.fresh ? 64 : 0) | (x.last ? 128 : 0));
    }
  }
}

static RecPlan plan_recurrent(const drnmf_handle* h, int B, int KS, int NB, int G, bool bwd) {
  RecPlan p{};
  p.ok = false;
  const int Rp = h->Rp;
  p.MT = Rp / 128;
  p.NB = NB;
  p.n_tiles_total = (B + p.NB - 1) / p.NB;
  if (G < 1) G = 1;
  if (G > p.n_tiles_total) G = p.n_tiles_total;
  p.n_tiles = (p.n_tiles_total + G - 1) / G;
  p.G = (p.n_tiles_total + p.n_tiles - 1) / p.n_tiles;      // no empty group
  if (Rp % (KS * 32) != 0 || p.MT * KS * p.G > h->num_sms) { p.why = "no (M-tile x K-split) grid fits the device"; return p; }
  p.KS = KS; p.RO = 128 / KS; p.KSLICE = Rp / KS; p.NSC = (p.KSLICE + 63) / 64;
  if (p.NSC > RT_MAXSC) { p.why = "K-slice wider than 1024 atoms"; return p; }
  {   // who publishes: every owner warp with its own red.release (default: measured faster at 1, 2 and 8 tiles per
      // group) or a publisher thread that batches the gpu-scope fences (DRNMF_REC_PUB=thread)
    const char* e = getenv("DRNMF_REC_PUB");
    p.a.pub_unit = (e && !strcmp(e, "thread")) ? 1 : 4;
  }
  // 128 owner threads x (4 rows x 4 batch columns) = 2048 outputs; the forward pass takes 4096 with the split epilogue
  p.a.split = (!bwd && p.NB == 64 && p.RO == 64 && !getenv("DRNMF_REC_NOSPLIT")) ? 1 : 0;
  const int ro_grp = p.a.split ? p.RO / 2 : p.RO;                      // rows per owner group
  if (ro_grp * p.NB > 2048 || p.RO % 8 != 0) { p.why = "rows per owner x batch tile exceeds the per-thread output budget"; return p; }
  if (p.a.split) p.a.pub_unit = 8;                                     // both groups publish: 8 flag increments per (CTA, item)
  p.CB = ro_grp * p.NB >= 2048 ? 4 : (ro_grp * p.NB >= 1024 ? 2 : 1);  // batch columns per owner thread (4 rows x CB)
  if (((ro_grp / 4) * (p.NB / p.CB)) % 32 != 0) { p.why = "owner tile smaller than a warp"; return p; }
  const int h_stage = 4 * p.NB * 128, red_slot = 128 * p.NB * 4;      // 64 atoms (hi, lo) x NB ; KS blocks of RO x NB fp32
  const int leak_b = round_up(2 * p.n_tiles * p.NB * 4, 128), out_b = round_up(max(p.NB * (p.RO + 2), 128) * 4, 128);
  const int fixed = leak_b + out_b + (int)sizeof(RecBars) + 256;
  const int budget = 232448 - 1024 - fixed;
  // every reduction slot has a twin staging slot on the pusher side (same index), hence 2 * red_slot per depth.
  // One batch tile per group (latency mode): the next step's hidden state only exists after this step's exchange, so a
  // step's worth of hidden sub-chunks and one reduction slot are enough and the rest of the shared memory holds the
  // weight ring (a whole step ahead when it fits).  Several tiles (throughput mode): the tiles pipeline, so the
  // reduction slots and the hidden-state ring get their second stages first.
  const int w_stage = 16384;
  const int env_wst = getenv("DRNMF_REC_WST") ? atoi(getenv("DRNMF_REC_WST")) : 0;
  const int env_hst = getenv("DRNMF_REC_HST") ? atoi(getenv("DRNMF_REC_HST")) : 0;
  const int env_rst = getenv("DRNMF_REC_RST") ? atoi(getenv("DRNMF_REC_RST")) : 0;
  const int pieces_per_step = p.KSLICE / 32;
  p.HST = 2; p.RST = 1; p.WST = 2;
  int rem = budget - p.HST * h_stage - p.RST * 2 * red_slot - p.WST * w_stage;
  if (rem < 0) { p.why = "hidden-state, reduction and weight rings do not fit in shared memory"; return p; }
  auto grow = [&](int& depth, int unit, int cap) { while (depth < cap && rem >= unit) { ++depth; rem -= unit; } };
  if (p.n_tiles == 1) {
    grow(p.HST, h_stage, min(p.NSC, RT_MAXH));
    grow(p.WST, w_stage, min(pieces_per_step, RT_MAXW));
  } else if (B <= 64) {
    // two pipelined tiles of the latency regime: a deeper weight ring only adds TMA traffic ahead of the hidden-state
    // loads that are on the critical path (measured 6.15 -> 6.00 us/step with 2 instead of 4 pieces in flight)
    grow(p.RST, 2 * red_slot, 2);
    grow(p.HST, h_stage, min(2 * p.NSC, RT_MAXH));
  } else if (p.a.split) {
    // K-split 2 x 64 columns: the loaders bound the item and every byte prefetched ahead competes with them for the
    // shared-memory port - the minimal rings are the fastest (measured: H2 W2 115 TF/s, H3 W2 112, H2 W4 113)
  } else {
    grow(p.RST, 2 * red_slot, 2);
    grow(p.HST, h_stage, 3);
    grow(p.WST, w_stage, 4);
    grow(p.HST, h_stage, min(2 * p.NSC, RT_MAXH));
    grow(p.WST, w_stage, min(pieces_per_step, RT_MAXW));
  }
  if (env_hst > 0 && env_hst <= RT_MAXH) p.HST = env_hst;
  if (env_rst > 0 && env_rst <= 4) p.RST = env_rst;
  if (env_wst > 0 && env_wst <= RT_MAXW) p.WST = env_wst;
  if (p.HST * h_stage + p.RST * 2 * red_slot + p.WST * w_stage > budget) { p.why = "requested ring depths do not fit in shared memory"; return p; }
  int off = 0;
  p.a.off_w = off; off += p.WST * w_stage;
  p.a.off_h = off; off += p.HST * h_stage;
  p.a.off_red = off; off += p.RST * red_slot;
  p.a.off_push = off; off += p.RST * red_slot;
  p.a.off_leak = off; off += leak_b;
  p.a.off_out = off; off += out_b;
  p.a.off_bar = off; off += (int)sizeof(RecBars);
  p.smem = (size_t)off + 1024;
  p.a.h_stage_bytes = h_stage; p.a.red_slot_bytes = red_slot;
  p.a.MT = p.MT; p.a.KS = p.KS; p.a.RO = p.RO; p.a.KSLICE = p.KSLICE; p.a.n_tiles = p.n_tiles; p.a.NSC = p.NSC;
  p.a.n_tiles_total = p.n_tiles_total;
  p.a.WST = p.WST; p.a.HST = p.HST; p.a.RST = p.RST;
  build_schedule(p.sch.e, p.NSC, p.n_tiles, &p.a.rot);
  build_schedule(p.sch.e_last, p.NSC, p.n_tiles_total - (p.G - 1) * p.n_tiles, &p.a.rot);
  // scalar alph: S_k is symmetric -> mirrored fetches below the diagonal (64-aligned sub-chunks).  Latency regime only:
  // there they keep the weights L2-resident (1.6 GB instead of 16.5 GB of DRAM traffic per launch, 3 % faster); with
  // several tiles / groups the weights are reused anyway and the transposed loader costs 4 - 10 %.
  p.a.sym = (h->alph_dim == 1 && p.KSLICE % 64 == 0 && p.n_tiles * p.G <= 2 && !p.a.split && !getenv("DRNMF_REC_NOSYM")) ? 1 : 0;
  p.ok = true;
  return p;
}

using RecKernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const RecSched, const RecArgs);

// SYM (mirrored weight fetches, scalar alph) is a template parameter: the transposed loader costs registers and issue
// slots that the throughput regime, whose loaders are on the critical path, should not pay for.
template <int NB, bool SYM>
static RecKernel rec_kernel_nb(bool bwd, int CB) {
  if (bwd) return CB == 4 ? k_recurrent_tc<NB, true, 4, SYM> : (CB == 2 ? k_recurrent_tc<NB, true, 2, SYM> : k_recurrent_tc<NB, true, 1, SYM>);
  return CB == 4 ? k_recurrent_tc<NB, false, 4, SYM> : (CB == 2 ? k_recurrent_tc<NB, false, 2, SYM> : k_recurrent_tc<NB, false, 1, SYM>);
}
static RecKernel rec_kernel(const RecPlan& p, bool bwd) {
  if (p.a.split) return k_recurrent_tc<64, false, 4, false, true>;     // the only shape with a split epilogue (see plan_recurrent)
  if (p.a.sym)
    return p.NB == 16 ? rec_kernel_nb<16, true>(bwd, p.CB) : (p.NB == 32 ? rec_kernel_nb<32, true>(bwd, p.CB) : rec_kernel_nb<64, true>(bwd, p.CB));
  return p.NB == 16 ? rec_kernel_nb<16, false>(bwd, p.CB) : (p.NB == 32 ? rec_kernel_nb<32, false>(bwd, p.CB) : rec_kernel_nb<64, false>(bwd, p.CB));
}

static void rec_launch_config(const RecPlan& p, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(p.KS, p.MT, p.G);
  cfg.blockDim = dim3(RT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.KS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // The CTAs spin on flags written by other clusters, so the whole grid must be co-resident.  A cooperative launch
  // makes the driver check that at launch time (SMs taken by another stream / process / NCCL kernel -> the launch
  // fails cleanly with cudaErrorCooperativeLaunchTooLarge instead of running into the device-side watchdog).
  // (Nsight Compute cannot replay cooperative launches of this kernel - the driver reports LaunchFailed -, so the
  //  attribute is dropped when its injection environment is present; DRNMF_REC_COOP=0/1 overrides either way.)
  static const bool coop = getenv("DRNMF_REC_COOP") ? strcmp(getenv("DRNMF_REC_COOP"), "0") != 0
                                                    : getenv("NV_NSIGHT_INJECTION_PORT_BASE") == nullptr;
  if (coop) {
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.numAttrs = 2;
  }
}

// co-resident clusters the device offers for this plan (the kernel spins on peers: all CTAs must be resident)
static int rec_max_clusters(const RecPlan& p, bool bwd, int* out) {
  RecKernel kern = rec_kernel(p, bwd);          // the instantiation that will be launched (register use differs)
  DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, p.KS > 8 ? 1 : 0));
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
  rec_launch_config(p, cfg, attr, nullptr);
  cfg.numAttrs = 1;                             // the occupancy query takes the cluster shape only
  *out = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(out, kern, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); *out = 0; }
  return DRNMF_OK;
}

static int launch_rec(const RecPlan& p, bool bwd, const CUtensorMap& tH_hi, const CUtensorMap& tH_lo, const CUtensorMap& tW,
                      const CUtensorMap& tW64, cudaStream_t st) {
  RecKernel kern = rec_kernel(p, bwd);
  DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, p.KS > 8 ? 1 : 0));
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
  rec_launch_config(p, cfg, attr, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tH_hi, tH_lo, tW, tW64, p.sch, p.a);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("persistent recurrence launch failed (%s): grid %d x %d x %d CTAs must be co-resident; SMs in use by another "
              "stream or process?", cudaGetErrorString(e), p.KS, p.MT, p.G);
    return DRNMF_ERR_CUDA;
  }
  count_launch();
  return DRNMF_OK;
}

// Candidate tilings (measured on B200, R = 1000, K = 25; profiles/r2_recurrence_sweep.txt):
//   * latency regime (B <= 64 utterances): every step is a chain of ~8 dependent hops (TMA, MMA, commit, DSMEM, reduce,
//     release, flag, TMA) of 0.5 - 2k cycles each, insensitive to the bytes moved.  K-splits of 8 (cluster of 8, one
//     batch group: only 15 such clusters are co-resident) keep the per-CTA weight stream and MMA burst shortest:
//     B <= 32 runs one 32-column tile with the self-validating exchange (4.9 us/step), B <= 64 pipelines two 32-column
//     tiles through the same weights (6.1 us/step for both).
//   * throughput regime (B > 64): K-splits of 4 (33 co-resident clusters of 4) and batch groups on disjoint SMs
//     (R = 1000: 4 groups x 32 CTAs = 128 SMs); a 128 x (Rp/4) block per CTA does twice the tensor work per exchanged
//     partial tile.  The batch tile is the widest one that still leaves a tile for every group.
// When the preferred cluster size does not fit (large R: MT clusters of 8 are not co-resident) the next one is tried.
static RecPlan choose_plan(const drnmf_handle* h, int B, bool bwd) {
  RecPlan p{};
  p.ok = false; p.why = "no candidate tiling";
  const char* env_ks = getenv("DRNMF_REC_KS");
  const char* env_nb = getenv("DRNMF_REC_NB");
  const char* env_g = getenv("DRNMF_REC_G");
  const bool verbose = getenv("DRNMF_REC_VERBOSE") != nullptr;
  const bool latency = B <= 64;
  // Forward pass with >= 12 tiles of 64 utterances: K-split 2 (a third of the split-K exchange per flop, split epilogue,
  // see the owners) - measured 108 against 85 TF/s useful at B = 2048, equal at B = 512 where a group holds one tile.
  static const int ks_lat[5] = {8, 4, 2, 1, 16}, ks_thr[5] = {4, 8, 2, 1, 16}, ks_big[5] = {2, 4, 8, 1, 16};
  const bool big = !latency && !bwd && (B + 63) / 64 >= 12 && !getenv("DRNMF_REC_NOSPLIT");
  for (int kq = 0; kq < 5 && !p.ok; ++kq) {
    const int KS = latency ? ks_lat[kq] : (big ? ks_big[kq] : ks_thr[kq]);
    if (env_ks && atoi(env_ks) != KS) continue;
    // co-resident clusters for this cluster size (probe with the smallest tile: shared memory is at the limit anyway)
    RecPlan probe = plan_recurrent(h, B, KS, 16, 1, bwd);
    if (!probe.ok) { p.why = probe.why; continue; }
    int mc = 0;
    rec_max_clusters(probe, bwd, &mc);
    if (verbose) fprintf(stderr, "[libdrnmf] plan probe KS=%d MT=%d smem=%zu: %d co-resident clusters\n", KS, probe.MT, probe.smem, mc);
    if (mc < probe.MT) { p.why = "not enough co-resident clusters for any tiling"; continue; }
    const int g_max = env_g ? atoi(env_g) : (latency ? 1 : mc / probe.MT);
    for (int NB = 64; NB >= 16 && !p.ok; NB >>= 1) {
      if (env_nb && atoi(env_nb) != NB) continue;
      if (!env_nb && latency && ((NB == 64) || (NB == 32 && B <= 16))) continue;       // 32-column tiles (16 for B <= 16)
      // throughput: the widest tile that still gives every group a tile; 32 columns is the floor (fewer groups then)
      if (!env_nb && !latency && NB == 16) continue;
      if (!env_nb && !latency && NB > 32 && (B + NB - 1) / NB < g_max) continue;
      for (int G = g_max; G >= 1 && !p.ok; --G) {
        RecPlan c = plan_recurrent(h, B, KS, NB, G, bwd);
        if (!c.ok) { p.why = c.why; continue; }
        if (c.G != G && G != g_max) continue;       // this group count was already tried
        int mc2 = 0;
        rec_max_clusters(c, bwd, &mc2);
        if (verbose) fprintf(stderr, "[libdrnmf]   candidate KS=%d NB=%d G=%d tiles/group=%d smem=%zu: %d co-resident clusters (need %d)\n", KS, NB, c.G, c.n_tiles, c.smem, mc2, c.MT * c.G);
        if (mc2 < c.MT * c.G) { p.why = "not enough co-resident clusters for any tiling"; continue; }
        p = c;
      }
    }
  }
  if (p.ok && p.n_tiles_total * p.MT > 16384) { p.ok = false; p.why = "too many batch tiles"; }
  return p;
}

// Hidden-state tensor maps: 3-D slab maps (one instruction per 64-atom sub-chunk and hi/lo) when the driver accepts
// the stride order, else 2-D maps (one instruction per 32-atom tile).
static int make_h_maps(CUtensorMap* hi, CUtensorMap* lo, const FwdWorkspace& w, int Rp, int NB, int* h3d) {
  static int use3d = getenv("DRNMF_REC_H2D") ? 0 : 1;
  if (use3d) {
    if (make_tmap_slabs(hi, w.hb_hi, Rp, 2ull * w.Bp, Rp, NB, 2) == DRNMF_OK &&
        make_tmap_slabs(lo, w.hb_lo, Rp, 2ull * w.Bp, Rp, NB, 2) == DRNMF_OK) { *h3d = 1; return DRNMF_OK; }
    fprintf(stderr, "[libdrnmf] 3-D slab tensor maps rejected (%s); using 2-D maps\n", last_error());
    use3d = 0;
  }
  *h3d = 0;
  int rc;
  if ((rc = make_tmap_2d(hi, w.hb_hi, Rp, 2ull * w.Bp, Rp, 32, NB))) return rc;
  return make_tmap_2d(lo, w.hb_lo, Rp, 2ull * w.Bp, Rp, 32, NB);
}

static void record_cfg(const RecPlan& p, int* cfg8, int* groups) {
  const int c[8] = {p.NB, p.KS, p.MT, p.NSC, p.n_tiles, p.WST, p.HST, p.RST};
  for (int i = 0; i < 8; ++i) cfg8[i] = c[i];
  *groups = p.G;
}

// Backward chain on the persistent kernel.  Returns 1 (error text set) when no tiling covers the shape.
int launch_recurrent_bwd_tc(drnmf_handle* h, FwdWorkspace& w, int B, int T, const float* dH, float* deltaT_hi,
                            float* deltaT_lo, float* G, float* psum2, cudaStream_t st, unsigned int* progress, bool* progress_ok) {
  const int K = h->K, Rp = h->Rp;
  if (K < 2) return 1;
  RecPlan p = choose_plan(h, B, true);
  if (!p.ok) { set_error("persistent tcgen05 backward chain unavailable for this shape (%s)", p.why); return 1; }
  record_cfg(p, h->bwd_cfg, &h->bwd_groups);
  RecArgs& a = p.a;
  a.XW = nullptr; a.mvalid = w.mvalid; a.h0 = h->h0;
  a.xw_tmajor = 0; a.xw_t0 = 0; a.xw_ready = nullptr; a.Btot = B; a.started = nullptr;
  const bool prog_ok = (p.n_tiles_total == 1 && p.G == 1);
  a.progress = prog_ok ? progress : nullptr;
  if (progress_ok) *progress_ok = prog_ok;
  a.state = nullptr; a.psum = nullptr; a.Hp_hi = nullptr; a.Hp_lo = nullptr; a.H_user = nullptr;
  a.hb_hi = w.hb_hi; a.hb_lo = w.hb_lo; a.flags = w.flags; a.dev_error = h->dev_error;
  a.actT_hi = w.actT_hi; a.actT_lo = w.actT_lo;
  a.dH = dH; a.deltaT_hi = deltaT_hi; a.deltaT_lo = deltaT_lo; a.G = G; a.psum2 = psum2;
  a.alph_vec = (h->alph_dim > 1) ? h->alph : nullptr;
  a.d0mo_b = h->u0_d - h->u0_o; a.o0_b = h->u0_o; a.ok_b = h->uk_o;
  a.dbg = nullptr;
  a.B = B; a.Bp = w.Bp; a.T = T; a.K = K; a.R = h->R; a.Rp = Rp;
  DRNMF_CHECK(p.n_tiles_total * p.NB <= w.Bp, "batch tiles (%d x %d) exceed the padded batch %d", p.n_tiles_total, p.NB, w.Bp);
  a.u0_dmo = 0.f; a.u0_off = 0.f; a.uk_dmo = 0.f; a.uk_off = 0.f;
  DRNMF_CUDA(cudaMemsetAsync(w.flags, 0, sizeof(unsigned int) * p.n_tiles_total * p.MT, st));
  // latency mode (one batch tile per group, NB <= 32): self-validating exchange; the ping-pong buffer starts zeroed
  // (measured: it wins with one or two groups - 5.3 -> 4.9 us/step at B = 32 - and loses when 128 CTAs pull through LDG)
  a.ll = (p.n_tiles <= (getenv("DRNMF_REC_LLT") ? atoi(getenv("DRNMF_REC_LLT")) : 1) && p.NB <= 32 && p.G <= 2 && a.pub_unit == 4 && !(getenv("DRNMF_REC_LL") && !strcmp(getenv("DRNMF_REC_LL"), "0"))) ? 1 : 0;
  if (a.ll) DRNMF_CUDA(cudaMemsetAsync(w.hb_hi, 0, sizeof(float) * 2 * (size_t)w.Bp * Rp, st));
  CUtensorMap tH_hi, tH_lo, tW, tW64;
  int rc;
  if ((rc = make_tmap_2d(&tW, h->ST_hi, Rp, (uint64_t)(K > 1 ? K - 1 : 1) * Rp, Rp, 32, 128))) return rc;
  if ((rc = make_tmap_2d(&tW64, h->ST_hi, Rp, (uint64_t)(K > 1 ? K - 1 : 1) * Rp, Rp, 32, 64))) return rc;
  if ((rc = make_h_maps(&tH_hi, &tH_lo, w, Rp, p.NB, &a.h3d))) return rc;
  return launch_rec(p, true, tH_hi, tH_lo, tW, tW64, st);
}

int recurrent_plan_ctas(const drnmf_handle* h, int B) {
  const RecPlan p = choose_plan(h, B, false);
  // (the pipelined order is compiled into the 16- / 32-column variants only: any other plan counts as "too large")
  return (p.ok && p.NB <= 32) ? p.KS * p.MT * p.G : (1 << 20);
}

int launch_recurrent_tc(drnmf_handle* h, FwdWorkspace& w, int B, int T, float* H_user, cudaStream_t st) {
  const int K = h->K, Rp = h->Rp;
  RecPlan p = choose_plan(h, B, false);
  if (!p.ok) {
    // No silent second backend: a shape the persistent kernel cannot tile is an error.  The CUDA-core recurrence only
    // runs when it is asked for (DRNMF_IMPL_SIMT handle, or DRNMF_RECURRENT=simt for debugging).
    set_error("persistent tcgen05 recurrence unavailable for R=%d (padded %d), B=%d: %s; the CUDA-core recurrence must be "
              "requested explicitly (DRNMF_IMPL_SIMT)", h->R, h->Rp, B, p.why);
    return DRNMF_ERR_INVALID;
  }
  h->last_rec_impl = 0;
  record_cfg(p, h->rec_cfg, &h->rec_groups);
  RecArgs& a = p.a;
  a.XW = w.XW; a.mvalid = w.mvalid; a.h0 = h->h0;
  DRNMF_CHECK(!w.xw_tmajor || p.NB <= 32, "pipelined forward needs a 16- or 32-column plan (got NB=%d)", p.NB);
  a.xw_tmajor = w.xw_tmajor; a.xw_t0 = w.xw_t0; a.xw_ready = w.xw_tmajor ? w.xw_ready : nullptr; a.Btot = B;
  a.started = w.xw_tmajor ? w.xw_ready + 1 : nullptr;
  a.progress = nullptr;
  a.state = w.state; a.psum = w.psum; a.Hp_hi = w.Hp_hi; a.Hp_lo = w.Hp_lo; a.H_user = H_user;
  a.hb_hi = w.hb_hi; a.hb_lo = w.hb_lo; a.flags = w.flags; a.dev_error = h->dev_error;
  a.actT_hi = w.actT_hi; a.actT_lo = w.actT_lo;
  static long long* dbg_dev = nullptr;
  const bool want_dbg = getenv("DRNMF_REC_DEBUG") != nullptr;
  if (want_dbg && !dbg_dev) DRNMF_CUDA(cudaMalloc(&dbg_dev, 16 * 8 * sizeof(long long)));
  a.dbg = want_dbg ? dbg_dev : nullptr;
  a.dbg_m = want_dbg ? atoi(getenv("DRNMF_REC_DEBUG")) - 1 : 0;      // DRNMF_REC_DEBUG=1 observes CTA (0,0), =2 CTA (0,1) ...
  if (a.dbg_m < 0 || a.dbg_m >= p.MT) a.dbg_m = 0;
  if (want_dbg) DRNMF_CUDA(cudaMemsetAsync(dbg_dev, 0, 16 * 8 * sizeof(long long), st));
  static long long* trace_dev = nullptr;
  const char* tr = want_dbg ? getenv("DRNMF_REC_TRACE") : nullptr;      // "lo:hi" = items of the observed CTA to trace
  a.trace = nullptr; a.trace_lo = a.trace_hi = 0;
  if (tr) {
    if (!trace_dev) DRNMF_CUDA(cudaMalloc(&trace_dev, (8 + 8 * RT_TRC_PER_ROLE * 2) * sizeof(long long)));
    DRNMF_CUDA(cudaMemsetAsync(trace_dev, 0, (8 + 8 * RT_TRC_PER_ROLE * 2) * sizeof(long long), st));
    a.trace = trace_dev;
    a.trace_lo = atoi(tr);
    a.trace_hi = strchr(tr, ':') ? atoi(strchr(tr, ':') + 1) : a.trace_lo + 2;
  }
  a.B = B; a.Bp = w.Bp; a.T = T; a.K = K; a.R = h->R; a.Rp = Rp;
  DRNMF_CHECK(p.n_tiles_total * p.NB <= w.Bp, "batch tiles (%d x %d) exceed the padded batch %d", p.n_tiles_total, p.NB, w.Bp);
  a.u0_dmo = h->u0_d - h->u0_o; a.u0_off = h->u0_o; a.uk_dmo = h->uk_d - h->uk_o; a.uk_off = h->uk_o;
  DRNMF_CUDA(cudaMemsetAsync(w.flags, 0, sizeof(unsigned int) * p.n_tiles_total * p.MT, st));
  // latency mode (one batch tile per group, NB <= 32): self-validating exchange; the ping-pong buffer starts zeroed
  // (measured: it wins with one or two groups - 5.3 -> 4.9 us/step at B = 32 - and loses when 128 CTAs pull through LDG)
  a.ll = (p.n_tiles <= (getenv("DRNMF_REC_LLT") ? atoi(getenv("DRNMF_REC_LLT")) : 1) && p.NB <= 32 && p.G <= 2 && a.pub_unit == 4 && !(getenv("DRNMF_REC_LL") && !strcmp(getenv("DRNMF_REC_LL"), "0"))) ? 1 : 0;
  if (a.ll) DRNMF_CUDA(cudaMemsetAsync(w.hb_hi, 0, sizeof(float) * 2 * (size_t)w.Bp * Rp, st));
  CUtensorMap tH_hi, tH_lo, tW, tW64;
  int rc;
  if ((rc = make_tmap_2d(&tW, h->ST_hi, Rp, (uint64_t)(K > 1 ? K - 1 : 1) * Rp, Rp, 32, 128))) return rc;
  if ((rc = make_tmap_2d(&tW64, h->ST_hi, Rp, (uint64_t)(K > 1 ? K - 1 : 1) * Rp, Rp, 32, 64))) return rc;
  if ((rc = make_h_maps(&tH_hi, &tH_lo, w, Rp, p.NB, &a.h3d))) return rc;
  rc = launch_rec(p, false, tH_hi, tH_lo, tW, tW64, st);
  if (rc == DRNMF_OK && want_dbg) {
    long long d[16 * 8];
    DRNMF_CUDA(cudaMemcpyAsync(d, dbg_dev, sizeof(d), cudaMemcpyDeviceToHost, st));
    DRNMF_CUDA(cudaStreamSynchronize(st));
    const char* names[16] = {"w-tma", "h-loader", "mma", "publisher", "pusher", "", "", "", "owner", "", "", "", "w-loader", "", "", ""};
    const long long items = (long long)T * (K - 1) * p.n_tiles;
    fprintf(stderr, "[libdrnmf] recurrence debug (CTA 0,0; cycles per MMA item, %lld items): NB=%d KS=%d G=%d tiles/group=%d NSC=%d W%d H%d R%d\n",
            items, p.NB, p.KS, p.G, p.n_tiles, p.NSC, p.WST, p.HST, p.RST);
    for (int wv = 0; wv < 16; ++wv) {
      if (!names[wv][0]) continue;
      const long long* q = d + wv * 8;
      fprintf(stderr, "  %-10s total %8.0f | acc0 %8.0f acc1 %8.0f acc2 %8.0f acc3 %8.0f acc4 %8.0f acc5 %8.0f acc6 %8.0f\n", names[wv],
              (double)q[0] / items, (double)q[1] / items, (double)q[2] / items, (double)q[3] / items, (double)q[4] / items,
              (double)q[5] / items, (double)q[6] / items, (double)q[7] / items);
    }
  }
  if (rc == DRNMF_OK && a.trace) {
    static long long tb[8 + 8 * RT_TRC_PER_ROLE * 2];
    DRNMF_CUDA(cudaMemcpyAsync(tb, trace_dev, sizeof(tb), cudaMemcpyDeviceToHost, st));
    DRNMF_CUDA(cudaStreamSynchronize(st));
    struct Ev { long long t; int role, ev, item; };
    static Ev evs[8 * RT_TRC_PER_ROLE];
    int n = 0;
    for (int r = 0; r < 8; ++r)
      for (int e = 0; e < tb[r] && e < RT_TRC_PER_ROLE; ++e) {
        const long long tag = tb[8 + (r * RT_TRC_PER_ROLE + e) * 2];
        evs[n++] = Ev{tb[8 + (r * RT_TRC_PER_ROLE + e) * 2 + 1], r, (int)(tag >> 16), (int)(tag & 0xFFFF)};
      }
    for (int i = 1; i < n; ++i)
      for (int j = i; j > 0 && evs[j].t < evs[j - 1].t; --j) { Ev x = evs[j]; evs[j] = evs[j - 1]; evs[j - 1] = x; }
    static const char* rn[8] = {"w-tma", "h-load", "mma", "pub", "pusher", "owner", "w-load", "?"};
    fprintf(stderr, "[libdrnmf] trace of CTA (0,%d), items [%d,%d): role event item +cycles\n", a.dbg_m, a.trace_lo, a.trace_hi);
    for (int i = 0; i < n; ++i)
      fprintf(stderr, "  %-7s ev%-3d item %-5d +%lld\n", rn[evs[i].role], evs[i].ev, evs[i].item, evs[i].t - evs[0].t);
  }
  return rc;
}

}  // namespace drnmf
