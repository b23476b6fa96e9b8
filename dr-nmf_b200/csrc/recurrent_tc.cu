// The DR-NMF recurrence (custom_layers.py:343-375 under Keras' masked scan) as ONE persistent sm_100a kernel.
//
// Problem shape: per utterance the chain is T x K_layers strictly serial steps (layer 0 of frame t+1 starts from the
// last layer of frame t); per step the work is  g^k = relu(g^{k-1} . S_k + x~W_k + b_k + leak),  a (B x R).(R x R)
// product with a SKINNY batch dimension.  S_k (4 MB at R=1000) must be re-streamed every step, so every SM has to pull
// its share of the weights each step: the step is tiled as
//       M-tile m (128 output atoms)  x  K-split s (a slice of the input atoms),   grid = (KS, MT), cluster = the KS
// K-splits of one M-tile.  Each CTA multiplies its 128 x Kslice block of S_k^T (A operand, TMA -> smem, 128B swizzle)
// by the Kslice x NB block of the hidden state (B operand) on the tensor cores (tcgen05.mma kind::tf32, 3xTF32
// compensation, fp32 accumulators in TMEM).  The split-K partial sums are reduced INSIDE the cluster through
// distributed shared memory (st.async + mbarrier complete_tx): CTA o of the cluster owns rows [o*RO, (o+1)*RO) of the
// M-tile, sums the KS partials in a fixed order (deterministic), applies the fused epilogue (input projection, bias,
// rank-1 leak, relu, Keras mask carry) and publishes hi/lo fp32 of the new hidden rows to a ping-pong global buffer
// (L2 resident) + a release flag.  Consumers acquire the flag and TMA the slice they need.  Utterances are cut into
// independent batch tiles of NB columns that are software-pipelined through the same weights, which hides the
// exchange latency when B is large and reuses every weight tile n_tiles times.
//
// Warp roles (384 threads): 0 weight TMA | 1 hidden-state TMA (+flag acquire) | 2 MMA issuer / TMEM owner | 3 idle |
//                           4-7 TMEM -> DSMEM pushers | 8-11 row owners (reduce + epilogue + publish).
#include "internal.h"

#include <cstdio>
#include <cstdlib>

namespace drnmf {

constexpr int RT_THREADS = 384;
constexpr int RT_AST = 2;                    // TMEM accumulator stages
constexpr long long RT_WATCHDOG = 3000000000LL;

struct RecArgs {
  // tensors
  const float* XW; const float* bias; const float* mvalid; const float* h0;
  float *state, *psum, *Hp_hi, *Hp_lo, *H_user, *hb_hi, *hb_lo;
  unsigned int* flags;
  int* dev_error;
  // shapes
  int B, Bp, T, K, R, Rp;
  int MT, KS, RO, ATOMS, KSLICE, n_tiles;
  int WST, HST, RST;                         // ring depths: weight atoms, hidden tiles, reduction slots
  float u0_dmo, u0_off, uk_dmo, uk_off;
  // smem offsets (bytes from the 1024-aligned base)
  int off_w, off_h, off_red, off_leak, off_out, off_bar;
  int h_stage_bytes, red_slot_bytes;
};

struct RecBars {   // all mbarriers, laid out at off_bar
  uint64_t w_full[8], w_empty[8], h_full[4], h_empty[4], t_full[RT_AST], t_empty[RT_AST], red_full[4], red_free[4];
  uint32_t tmem_slot;
  int abort;
};

// named-barrier AND-reduction over the 128 owner threads (barrier id 1): uniform agreement on a predicate
__device__ __forceinline__ bool owners_all(bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.and.pred p, 1, 128, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(r) : "r"((uint32_t)pred) : "memory");
  return r != 0;
}

__device__ __forceinline__ bool poll_flag(const unsigned int* f, unsigned int target, volatile int* err) {
  if (flag_ld_acquire(f) >= target) return true;
  long long t0 = clock64();
  unsigned it = 0;
  while (flag_ld_acquire(f) < target) {
    if ((++it & 0xFF) == 0) {
      if (clock64() - t0 > RT_WATCHDOG) return false;
      if (*err) return false;
    }
  }
  return true;
}

template <int NB>
__global__ void __launch_bounds__(RT_THREADS, 1)
k_recurrent_tc(const __grid_constant__ CUtensorMap tmS_hi, const __grid_constant__ CUtensorMap tmS_lo,
               const __grid_constant__ CUtensorMap tmH_hi, const __grid_constant__ CUtensorMap tmH_lo, RecArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  RecBars* bars = reinterpret_cast<RecBars*>(smem + a.off_bar);
  float* leak_s = reinterpret_cast<float*>(smem + a.off_leak);       // n_tiles x NB : sum_j state[b][j] of this frame
  float* out_s = reinterpret_cast<float*>(smem + a.off_out);         // NB x (RO+1) staging of the new state rows
  volatile int* err = a.dev_error;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x;            // K-split == rank in cluster
  const int m = blockIdx.y;            // M-tile
  const int K = a.K, T = a.T, Rp = a.Rp, n_tiles = a.n_tiles, ATOMS = a.ATOMS;
  constexpr int RSTRIDE = NB + 4;      // padded row of a reduction slot (bank-conflict-free transposed reads)
  constexpr uint32_t TMEM_COLS = (RT_AST * NB < 32) ? 32 : RT_AST * NB;
  const uint32_t W_ATOM_BYTES = 128 * 128;                           // 128 rows x 32 fp32
  const uint32_t H_ATOM_BYTES = NB * 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmS_hi); tma_prefetch_desc(&tmS_lo); tma_prefetch_desc(&tmH_hi); tma_prefetch_desc(&tmH_lo);
    for (int i = 0; i < 8; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&bars->h_full[i], 1); mbar_init(&bars->h_empty[i], 1); }
    for (int i = 0; i < RT_AST; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], 4); }
    for (int i = 0; i < 4; ++i) { mbar_init(&bars->red_full[i], 1); mbar_init(&bars->red_free[i], a.KS); }
    bars->abort = 0;
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(&bars->tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // every CTA's barriers exist before any remote arrive / st.async
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const int n_mma_steps = T * (K - 1);

  if (warp == 0) {
    // ================= weight producer: S_k^T[m*128.., s*KSLICE + a*32..] hi/lo, one atom per stage =================
    if (lane == 0) {
      int wl = 0;
      for (int t = 0; t < T && !*err; ++t)
        for (int k = 1; k < K; ++k)
          for (int at = 0; at < ATOMS; ++at, ++wl) {
            const int ws = wl % a.WST;
            if (!mbar_wait(&bars->w_empty[ws], ((wl / a.WST) & 1) ^ 1, err, RT_WATCHDOG)) { atomicCAS(a.dev_error, 0, 201); goto w_done; }
            uint8_t* dst = smem + a.off_w + ws * (2 * W_ATOM_BYTES);
            mbar_expect_tx(&bars->w_full[ws], 2 * W_ATOM_BYTES);
            const int c0 = s * a.KSLICE + at * 32, c1 = (k - 1) * Rp + m * 128;
            tma_load_2d(dst, &tmS_hi, &bars->w_full[ws], c0, c1);
            tma_load_2d(dst + W_ATOM_BYTES, &tmS_lo, &bars->w_full[ws], c0, c1);
          }
    }
  w_done:;
  } else if (warp == 1) {
    // ================= hidden-state loader: acquire the producers' flag, then TMA the K-slice of tile i =================
    if (lane == 0) {
      const int m_lo = (s * a.KSLICE) / 128, m_hi = ((s + 1) * a.KSLICE - 1) / 128;
      int it = 0;
      for (int t = 0; t < T; ++t)
        for (int k = 1; k < K; ++k) {
          const unsigned int target = (unsigned int)a.KS * (unsigned int)(t * K + k);   // step (t,k-1) published
          const int slot = (k - 1) & 1;
          for (int i = 0; i < n_tiles; ++i, ++it) {
            const int hs = it % a.HST;
            if (!mbar_wait(&bars->h_empty[hs], ((it / a.HST) & 1) ^ 1, err, RT_WATCHDOG)) { atomicCAS(a.dev_error, 0, 202); goto h_done; }
            for (int mm = m_lo; mm <= m_hi; ++mm)
              if (!poll_flag(a.flags + i * a.MT + mm, target, err)) { atomicCAS(a.dev_error, 0, 203); goto h_done; }
            fence_proxy_async();                       // generic-proxy writes of the owners -> async-proxy (TMA) reads
            uint8_t* dst = smem + a.off_h + hs * a.h_stage_bytes;
            mbar_expect_tx(&bars->h_full[hs], 2 * ATOMS * H_ATOM_BYTES);
            for (int at = 0; at < ATOMS; ++at) {
              const int c0 = s * a.KSLICE + at * 32, c1 = slot * a.Bp + i * NB;
              tma_load_2d(dst + (2 * at) * H_ATOM_BYTES, &tmH_hi, &bars->h_full[hs], c0, c1);
              tma_load_2d(dst + (2 * at + 1) * H_ATOM_BYTES, &tmH_lo, &bars->h_full[hs], c0, c1);
            }
          }
        }
    }
  h_done:;
  } else if (warp == 2) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, NB);
      int it = 0;
      for (int ms = 0; ms < n_mma_steps; ++ms) {
        for (int i = 0; i < n_tiles; ++i, ++it) {
          const int as = it % RT_AST, hs = it % a.HST;
          if (!mbar_wait(&bars->t_empty[as], ((it / RT_AST) & 1) ^ 1, err, RT_WATCHDOG)) { atomicCAS(a.dev_error, 0, 204); goto m_done; }
          if (!mbar_wait(&bars->h_full[hs], (it / a.HST) & 1, err, RT_WATCHDOG)) { atomicCAS(a.dev_error, 0, 205); goto m_done; }
          const uint32_t d_tmem = tmem_base + as * NB;
          const uint32_t hbase = smem_u32(smem + a.off_h + hs * a.h_stage_bytes);
          for (int at = 0; at < ATOMS; ++at) {
            const int wl = ms * ATOMS + at, ws = wl % a.WST;
            if (i == 0 && !mbar_wait(&bars->w_full[ws], (wl / a.WST) & 1, err, RT_WATCHDOG)) { atomicCAS(a.dev_error, 0, 206); goto m_done; }
            tc_fence_after();
            const uint32_t wbase = smem_u32(smem + a.off_w + ws * (2 * W_ATOM_BYTES));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t w_hi = umma_desc_k128(wbase + ks * 32);
              const uint64_t w_lo = umma_desc_k128(wbase + W_ATOM_BYTES + ks * 32);
              const uint64_t h_hi = umma_desc_k128(hbase + (2 * at) * H_ATOM_BYTES + ks * 32);
              const uint64_t h_lo = umma_desc_k128(hbase + (2 * at + 1) * H_ATOM_BYTES + ks * 32);
              umma_tf32(d_tmem, w_lo, h_hi, idesc, !(at == 0 && ks == 0));
              umma_tf32(d_tmem, w_hi, h_lo, idesc, true);
              umma_tf32(d_tmem, w_hi, h_hi, idesc, true);
            }
            if (i == n_tiles - 1) tc_commit(&bars->w_empty[ws]);     // weights of this step fully consumed
          }
          tc_commit(&bars->h_empty[hs]);
          tc_commit(&bars->t_full[as]);
        }
      }
    }
  m_done:;
  } else if (warp >= 4 && warp < 8) {
    // ================= pushers: TMEM accumulator rows -> owner CTA's reduction slot over DSMEM =================
    const int q = warp - 4;
    const int rho = q * 32 + lane;                   // accumulator row (TMEM lane) within the M-tile
    const int o = rho / a.RO, r = rho % a.RO;        // owner CTA in the cluster, row within the owner
    const uint32_t rem_red = mapa_u32(smem_u32(smem + a.off_red), (uint32_t)o);
    const uint32_t rem_bar = mapa_u32(smem_u32(&bars->red_full[0]), (uint32_t)o);
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    bool ok = true;
    for (int ms = 0; ms < n_mma_steps && ok; ++ms) {
      for (int i = 0; i < n_tiles; ++i, ++it) {
        const int as = it % RT_AST, rs = it % a.RST;
        if (!mbar_wait(&bars->t_full[as], (it / RT_AST) & 1, err, RT_WATCHDOG)) { atomicCAS(a.dev_error, 0, 207); ok = false; break; }
        tc_fence_after();
        float v[NB];
#pragma unroll
        for (int c = 0; c < NB; c += 16) tmem_ld16(trow + as * NB + c, v + c);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->t_empty[as]);
        if (!mbar_wait_cluster(&bars->red_free[rs], ((it / a.RST) & 1) ^ 1, err, RT_WATCHDOG)) { atomicCAS(a.dev_error, 0, 208); ok = false; break; }
        const uint32_t dst = rem_red + rs * a.red_slot_bytes + ((s * a.RO + r) * RSTRIDE) * 4;
        const uint32_t bar = rem_bar + rs * 8;
#pragma unroll
        for (int c = 0; c < NB; c += 4) st_async_v4(dst + c * 4, bar, v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
    }
  } else if (warp >= 8) {
    // ================= owners: reduce the KS partials of rows [o*RO, (o+1)*RO), epilogue, publish =================
    const int otid = threadIdx.x - 256;              // 0..127
    const int RO = a.RO;
    const int row0 = m * 128 + s * RO;               // first global output row this CTA owns
    const int n_out = RO * NB;                       // outputs per tile
    const int cta_lin = m * a.KS + s, n_cta = a.MT * a.KS;
    const size_t KRp = (size_t)K * Rp;
    // sum_j h0[j]: leak of frame 0 (state = h0 for every utterance), same fixed order in every CTA
    float h0sum = 0.f;
    for (int j = 0; j < a.R; ++j) h0sum += a.h0[j];
    int it = 0;
    bool ok = true;
    for (int t = 0; t < T && ok; ++t) {
      for (int k = 0; k < K && ok; ++k) {
        const bool last = (k == K - 1);
        const float dmo = (k == 0) ? a.u0_dmo : a.uk_dmo, off = (k == 0) ? a.u0_off : a.uk_off;
        for (int i = 0; i < n_tiles; ++i) {
          // ---- prefetch what does not depend on the exchange: x~W_k, bias, validity ----
          constexpr int MAXE = 8;                    // outputs per thread (RO*NB/128 <= 8 for the supported configs)
          float xw[MAXE], acc[MAXE];
#pragma unroll
          for (int c = 0; c < MAXE; ++c) {
            const int e = otid + 128 * c;
            xw[c] = 0.f; acc[c] = 0.f;
            if (e < n_out) {
              const int b = i * NB + e / RO, row = row0 + e % RO;
              if (b < a.B && row < a.R)
                xw[c] = __ldg(a.XW + ((size_t)b * T + t) * KRp + (size_t)k * Rp + row) + __ldg(a.bias + (size_t)k * Rp + row);
            }
          }
          if (k == 0) {
            // ---- frame start: leak[b] = sum_j state[b][j] from the published partial sums of the previous frame ----
            if (t > 0) {
              const unsigned int target = (unsigned int)a.KS * (unsigned int)(t * K);
              if (otid < a.MT && !poll_flag(a.flags + i * a.MT + otid, target, err)) { atomicCAS(a.dev_error, 0, 209); bars->abort = 1; }
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (bars->abort) { ok = false; break; }
              const int b = otid % NB, part = otid / NB, nparts = 128 / NB;
              float sacc = 0.f;
              const float* ps = a.psum + (size_t)((t - 1) & 1) * 256 * a.Bp + i * NB + b;
              for (int c = part; c < n_cta; c += nparts) sacc += __ldcg(ps + (size_t)c * a.Bp);
              out_s[part * NB + b] = sacc;
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (otid < NB) {
                float tot = 0.f;
                for (int p = 0; p < nparts; ++p) tot += out_s[p * NB + otid];
                leak_s[i * NB + otid] = tot;
              }
              asm volatile("bar.sync 1, 128;" ::: "memory");
            } else if (otid < NB) {
              leak_s[i * NB + otid] = h0sum;
            }
            if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          } else {
            // ---- wait for the KS partial tiles, sum them in rank order ----
            const int rs = it % a.RST;
            if (otid == 0) mbar_expect_tx(&bars->red_full[rs], (uint32_t)(128 * NB * 4));
            const bool got = mbar_wait_cluster(&bars->red_full[rs], (it / a.RST) & 1, err, RT_WATCHDOG);
            if (!owners_all(got)) { atomicCAS(a.dev_error, 0, 210); ok = false; break; }
            const float* red = reinterpret_cast<const float*>(smem + a.off_red + rs * a.red_slot_bytes);
#pragma unroll
            for (int c = 0; c < MAXE; ++c) {
              const int e = otid + 128 * c;
              if (e < n_out) {
                const int bl = e / RO, r = e % RO;
                float sum = 0.f;
                for (int src = 0; src < a.KS; ++src) sum += red[(src * RO + r) * RSTRIDE + bl];
                acc[c] = sum;
              }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");            // every owner thread is done reading the slot
            if (otid < a.KS) mbar_arrive_remote(&bars->red_free[rs], (uint32_t)otid);
            ++it;
          }
          // ---- fused epilogue ----
#pragma unroll
          for (int c = 0; c < MAXE; ++c) {
            const int e = otid + 128 * c;
            if (e < n_out) {
              const int bl = e / RO, r = e % RO, b = i * NB + bl, row = row0 + r;
              float g = 0.f, st_new = 0.f;
              const bool valid = (b < a.B && row < a.R);
              float st_old = 0.f;
              const bool need_state = valid && (k == 0 || dmo != 0.f);
              if (need_state) st_old = (t == 0) ? __ldg(a.h0 + row) : __ldcg(a.state + (size_t)b * Rp + row);
              if (valid) g = fmaxf(acc[c] + xw[c] + off * leak_s[i * NB + bl] + (need_state ? dmo * st_old : 0.f), 0.f);
              if (!last) {
                if (b < a.Bp) {
                  const size_t o2 = ((size_t)(k & 1) * a.Bp + b) * Rp + row;
                  __stcg(a.hb_hi + o2, g);
                  __stcg(a.hb_lo + o2, tf32_lo(g));
                }
              } else if (b < a.B) {
                // Keras masked scan: out_t = m ? g : out_{t-1} (zeros before the first step); state = m ? g : state
                const size_t bt = (size_t)b * T + t;
                const bool mv = __ldg(a.mvalid + bt) != 0.f;
                float outv;
                if (mv) { outv = g; st_new = g; }
                else {
                  outv = (t > 0) ? __ldcg(a.Hp_hi + (bt - 1) * Rp + row) : 0.f;
                  st_new = (row < a.R) ? ((t == 0) ? __ldg(a.h0 + row) : __ldcg(a.state + (size_t)b * Rp + row)) : 0.f;
                }
                __stcg(a.Hp_hi + bt * Rp + row, outv);
                __stcg(a.Hp_lo + bt * Rp + row, tf32_lo(outv));
                if (a.H_user && row < a.R) a.H_user[bt * a.R + row] = outv;
                __stcg(a.state + (size_t)b * Rp + row, st_new);
              }
              if (last) out_s[bl * (RO + 1) + r] = st_new;
            }
          }
          if (last) {
            // partial row sums of the new state over this CTA's RO rows (rank-1 leak of the next frame)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (otid < NB) {
              float ps = 0.f;
              for (int r = 0; r < RO; ++r) ps += out_s[otid * (RO + 1) + r];
              __stcg(a.psum + (size_t)(t & 1) * 256 * a.Bp + (size_t)cta_lin * a.Bp + i * NB + otid, ps);
            }
          }
          // ---- publish: all owner threads' global writes -> one release increment of flag[tile][m] ----
          fence_proxy_async();
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (otid == 0) flag_add_release(a.flags + i * a.MT + m, 1u);
        }
      }
    }
  }
  // ---- teardown: nobody leaves while a peer may still touch its shared memory ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<TMEM_COLS>(tmem_base); }
}

// ---------------------------------------------------------------------------------------------------
struct RecPlan { int NB, KS, MT, RO, ATOMS, KSLICE, n_tiles, WST, HST, RST; size_t smem; RecArgs a; bool ok; const char* why; };

static RecPlan plan_recurrent(const drnmf_handle* h, int B, int KS, int NB) {
  RecPlan p{};
  p.ok = false;
  const int Rp = h->Rp;
  p.MT = Rp / 128;
  p.NB = NB;
  p.n_tiles = (B + p.NB - 1) / p.NB;
  if (Rp % (KS * 32) != 0 || p.MT * KS > h->num_sms) { p.why = "no (M-tile x K-split) grid fits the device"; return p; }
  p.KS = KS; p.RO = 128 / KS; p.KSLICE = Rp / KS; p.ATOMS = p.KSLICE / 32;
  if (p.RO * p.NB > 128 * 8) { p.why = "rows per owner x batch tile exceeds the per-thread output budget"; return p; }
  const int w_atom = 2 * 128 * 128, h_stage = 2 * p.ATOMS * p.NB * 128, red_slot = 128 * (p.NB + 4) * 4;
  const int leak_b = round_up(p.n_tiles * p.NB * 4, 128), out_b = round_up(max(p.NB * (p.RO + 1), 128) * 4, 128);
  const int fixed = leak_b + out_b + (int)sizeof(RecBars) + 256;
  const int budget = 232448 - 1024 - fixed;
  p.HST = 2; p.RST = 2;
  int wst = (budget - p.HST * h_stage - p.RST * red_slot) / w_atom;
  if (wst > 8) wst = 8;
  if (wst > 2 * p.ATOMS) wst = 2 * p.ATOMS;
  if (wst < p.ATOMS + 1 && wst < 2 * p.ATOMS) {
    if (wst < p.ATOMS) { p.why = "K-slice of the weights does not fit in shared memory (needs the streaming variant)"; return p; }
  }
  p.WST = wst;
  int rem = budget - p.WST * w_atom - p.HST * h_stage - p.RST * red_slot;
  while (p.HST < 4 && rem >= h_stage) { ++p.HST; rem -= h_stage; if (p.HST >= 3) break; }
  while (p.RST < 4 && rem >= red_slot) { ++p.RST; rem -= red_slot; if (p.RST >= 3) break; }
  int off = 0;
  p.a.off_w = off; off += p.WST * w_atom;
  p.a.off_h = off; off += p.HST * h_stage;
  p.a.off_red = off; off += p.RST * red_slot;
  p.a.off_leak = off; off += leak_b;
  p.a.off_out = off; off += out_b;
  p.a.off_bar = off; off += (int)sizeof(RecBars);
  p.smem = (size_t)off + 1024;
  p.a.h_stage_bytes = h_stage; p.a.red_slot_bytes = red_slot;
  p.a.MT = p.MT; p.a.KS = p.KS; p.a.RO = p.RO; p.a.ATOMS = p.ATOMS; p.a.KSLICE = p.KSLICE; p.a.n_tiles = p.n_tiles;
  p.a.WST = p.WST; p.a.HST = p.HST; p.a.RST = p.RST;
  p.ok = true;
  return p;
}

// co-resident clusters the device offers for this plan (the kernel spins on peers: all CTAs must be resident)
template <int NB>
static int rec_max_clusters(const RecPlan& p, int* out) {
  auto kern = k_recurrent_tc<NB>;
  DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, p.KS > 8 ? 1 : 0));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.KS, p.MT, 1);
  cfg.blockDim = dim3(RT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = p.smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.KS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  *out = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(out, kern, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); *out = 0; }
  return DRNMF_OK;
}

template <int NB>
static int launch_rec(const RecPlan& p, const CUtensorMap& tS_hi, const CUtensorMap& tS_lo, const CUtensorMap& tH_hi,
                      const CUtensorMap& tH_lo, cudaStream_t st) {
  auto kern = k_recurrent_tc<NB>;
  DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, p.KS > 8 ? 1 : 0));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.KS, p.MT, 1);
  cfg.blockDim = dim3(RT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.KS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  DRNMF_CUDA(cudaLaunchKernelEx(&cfg, kern, tS_hi, tS_lo, tH_hi, tH_lo, p.a));
  count_launch();
  return DRNMF_OK;
}

int launch_recurrent_tc(drnmf_handle* h, FwdWorkspace& w, int B, int T, float* H_user, cudaStream_t st) {
  const int K = h->K, Rp = h->Rp;
  // Candidate tilings, preferred first: more K-splits = more SMs streaming the weights; the cluster (= the K-splits of
  // one M-tile) must be co-resident MT times, which depends on the board's GPC layout -> ask the occupancy API.
  RecPlan p{};
  p.ok = false; p.why = "no candidate tiling";
  const char* env_ks = getenv("DRNMF_REC_KS");
  const char* env_nb = getenv("DRNMF_REC_NB");
  for (int KS = 16; KS >= 1 && !p.ok; KS >>= 1) {
    if (env_ks && atoi(env_ks) != KS) continue;
    for (int NB = 32; NB >= 16 && !p.ok; NB >>= 1) {
      if (env_nb && atoi(env_nb) != NB) continue;
      if (NB == 32 && B <= 16) continue;
      RecPlan c = plan_recurrent(h, B, KS, NB);
      if (!c.ok) { if (!p.why || !p.ok) p.why = c.why; continue; }
      int mc = 0;
      if (NB == 16) rec_max_clusters<16>(c, &mc); else rec_max_clusters<32>(c, &mc);
      if (mc < c.MT) { p.why = "not enough co-resident clusters for any tiling"; continue; }
      p = c;
    }
  }
  if (!p.ok || p.n_tiles * p.MT > 16384) {
    // Shapes the persistent kernel does not cover yet run on the CUDA-core recurrence (still on the GPU).
    static bool warned = false;
    if (!warned) { fprintf(stderr, "[libdrnmf] persistent tcgen05 recurrence unavailable (%s); using the SIMT recurrence\n", p.ok ? "too many tiles" : p.why); warned = true; }
    h->last_rec_impl = 1;
    return launch_recurrent_simt(h, w, B, T, H_user, st);
  }
  h->last_rec_impl = 0;
  { int c[8] = {p.NB, p.KS, p.MT, p.ATOMS, p.n_tiles, p.WST, p.HST, p.RST}; for (int i = 0; i < 8; ++i) h->rec_cfg[i] = c[i]; }
  RecArgs& a = p.a;
  a.XW = w.XW; a.bias = h->bias; a.mvalid = w.mvalid; a.h0 = h->h0;
  a.state = w.state; a.psum = w.psum; a.Hp_hi = w.Hp_hi; a.Hp_lo = w.Hp_lo; a.H_user = H_user;
  a.hb_hi = w.hb_hi; a.hb_lo = w.hb_lo; a.flags = w.flags; a.dev_error = h->dev_error;
  a.B = B; a.Bp = w.Bp; a.T = T; a.K = K; a.R = h->R; a.Rp = Rp;
  a.u0_dmo = h->u0_d - h->u0_o; a.u0_off = h->u0_o; a.uk_dmo = h->uk_d - h->uk_o; a.uk_off = h->uk_o;
  DRNMF_CUDA(cudaMemsetAsync(w.flags, 0, sizeof(unsigned int) * p.n_tiles * p.MT, st));
  CUtensorMap tS_hi, tS_lo, tH_hi, tH_lo;
  int rc;
  const uint64_t s_rows = (uint64_t)(K > 1 ? K - 1 : 1) * Rp;
  if ((rc = make_tmap_2d(&tS_hi, h->ST_hi, Rp, s_rows, Rp, 32, 128))) return rc;
  if ((rc = make_tmap_2d(&tS_lo, h->ST_lo, Rp, s_rows, Rp, 32, 128))) return rc;
  if ((rc = make_tmap_2d(&tH_hi, w.hb_hi, Rp, 2ull * w.Bp, Rp, 32, p.NB))) return rc;
  if ((rc = make_tmap_2d(&tH_lo, w.hb_lo, Rp, 2ull * w.Bp, Rp, 32, p.NB))) return rc;
  if (p.NB == 16) return launch_rec<16>(p, tS_hi, tS_lo, tH_hi, tH_lo, st);
  return launch_rec<32>(p, tS_hi, tS_lo, tH_hi, tH_lo, st);
}

}  // namespace drnmf
