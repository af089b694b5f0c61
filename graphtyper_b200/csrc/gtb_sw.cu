// gtb_sw.cu -- discovery re-alignment kernel: B200 replacement of paw::pairwise_alignment as
// graphtyper's realign_to_indels calls it (src/typer/caller.cpp:1864-1870,2007; paw/include/paw/align/
// pairwise_alignment.hpp:146-388): semi-global affine-gap alignment of a read (query, global) against a haplotype
// window (database, both ends free), match +1 / mismatch -4 / gap open 7 (first gap base) / extend 1, followed by the
// soft-clipping pass (penalty 5) and database begin/end extraction along the traceback
// (alignment_results.hpp:392-447,451-669).  Results are bit-identical to paw: score, database_begin/end, clip_begin/end.
//
// One warp per (read, window) pair.  The DP runs as a 32-lane systolic wavefront: lane l owns query columns
// c*l+1 .. c*l+c (c = ceil(read length / 32) <= 5, one template instance each), at step s it computes database row
// s-l+1 and hands (H', E, H) of its last column to lane l+1 with three shuffles.  A cell is two fused add-max DPX
// operations (VIADDMNMX: E and F), two VIMNMX and four funnel shifts that push the sign of (loser - winner) into the
// row word -- exactly paw's strict-greater backtrack bits (ins_extend, del_extend, ins, del), no predicates, no
// branches.  A lane's cells of one row are one 32-bit word; a row is one coalesced 128-byte store into the warp's
// scratch slab.  The traceback (clipping + begin/end in one walk) is executed redundantly by all lanes on rows staged
// 16 at a time in shared memory -- one bulk asynchronous copy (cp.async.bulk, the 1-D form of TMA: UBLKCP in SASS) of up
// to 2 KiB per refill, completed on a per-buffer mbarrier; double-buffered so the next 16 rows are in flight while the
// current ones are walked (sw_kernel<true>; sw_kernel<false> keeps the per-lane cp.async staging, GTB_SW_BULK=0).  The resident-warp count is sized so that all slabs stay in L2 (sw_resident_warps below).
// No tensor cores: this is min/max/add dynamic programming, not a contraction.

#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "gtb_device.cuh"

namespace gtb
{
constexpr int SW_CPL = 5;                 // query columns per lane
constexpr int SW_MAX_Q = 32 * SW_CPL;     // 160 >= MAX_READ_LENGTH (151)
constexpr int SW_WARPS = 4;               // warps per block
constexpr int SW_MIN_BLOCKS = 8;          // 32 resident warps / SM
constexpr int SW_ROWS_CACHED = 16;        // backtrack rows staged per refill
constexpr int SW_NEG = -(1 << 28);

namespace
{
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int MATCH = 1, MISMATCH = 4, GO = 7, GE = 1, CLIP = 5;

__device__ __forceinline__ bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
} // namespace

// DP phase with CPL query columns per lane (CPL = ceil(m / 32) keeps every lane busy for short reads).  Writes one
// backtrack word per (row, lane) and returns H[n][m] to every lane.
template <int CPL>
__device__ __forceinline__ int sw_dp(const uint8_t * q, int m, int n, const uint8_t * sdb, uint32_t * bt, int lane)
{
  int const c0 = lane * CPL; // columns c0+1 .. c0+CPL
  // Columns past the read end carry a sentinel base that matches nothing: they are computed (no per-column branch)
  // but never read -- the score comes from column m and the traceback never moves right of it.
  int qc[CPL], gop[CPL];
  int Hprev[CPL], Fprev[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c)
  {
    int const j = c0 + c + 1;
    // a query base matches the database base iff it is an upper-case A/C/G/T and the database base is the same letter
    // in either case (alignment_cache.hpp:70-121, libsimdpp_utils.hpp:96-123); everything else matches nothing
    qc[c] = (j <= m && is_acgt(q[j - 1])) ? (int)(q[j - 1] | 0x20) : 0x1FE;
    gop[c] = j == m ? 0 : GO;        // right_column_free: trailing database bases cost nothing
    Hprev[c] = -(GO + (j - 1) * GE); // initial row (alignment_options.hpp:282-299)
    Fprev[c] = SW_NEG;
  }
  int Hleft_prev = lane == 0 ? 0 : -(GO + (c0 - 1) * GE); // H[0][c0]
  int pub_Hp = 0, pub_E = SW_NEG, pub_H = 0;
  int const last_lane = (m - 1) / CPL;
  int const steps = n + last_lane;
  bool const lane_on = c0 < m;
  // database base of the row handled at the next step, fetched one step ahead of its use
  int dc_next = (lane == 0) ? sdb[0] | 0x20 : 0;
#pragma unroll 1 // five template variants + the traceback share the instruction cache: keep every loop body single
  for (int s = 0; s < steps; ++s)
  {
    int rHp = __shfl_up_sync(FULL, pub_Hp, 1);
    int rE = __shfl_up_sync(FULL, pub_E, 1);
    int rH = __shfl_up_sync(FULL, pub_H, 1);
    if (lane == 0) // column 0: leading database bases are free (left_column_free)
    {
      rHp = 0;
      rE = SW_NEG;
      rH = 0;
    }
    int const i = s - lane + 1;
    int const dc = dc_next;
    dc_next = sdb[min(max(i, 0), n - 1)] | 0x20; // row i + 1
    if (i >= 1 && i <= n && lane_on)
    {
      int left_Hp = rHp, left_E = rE, diag_src = Hleft_prev;
      // Backtrack bits are shifted in one at a time: (word << 1) | sign(difference) is a single funnel shift.  Per
      // cell the order ins_extend, del_extend, ins, del reproduces paw's nibble (8,4,2,1); column c of this lane
      // ends up in bits [4*(CPL-1-c), 4*(CPL-1-c)+3].  The strict-greater tie rules are paw's.
      uint32_t word = 0;
#pragma unroll
      for (int c = 0; c < CPL; ++c)
      {
        int const diag = diag_src + (qc[c] == dc ? MATCH : -MISMATCH);
        int const hup = Hprev[c];
        int const fopen = hup - gop[c];
        int const eopen = left_Hp - GO;
        // F = max(F_up - extend, open): extension strictly better <=> open - F_up + extend < 0
        word = __funnelshift_l((uint32_t)(fopen - Fprev[c] + GE), word, 1);
        int const f = __viaddmax_s32(Fprev[c], -GE, fopen);
        word = __funnelshift_l((uint32_t)(eopen - left_E + GE), word, 1);
        int const e = __viaddmax_s32(left_E, -GE, eopen);
        word = __funnelshift_l((uint32_t)(diag - f), word, 1); // ins: f > diag
        int const hp = max(diag, f);
        word = __funnelshift_l((uint32_t)(hp - e), word, 1); // del: e > hp
        int const h = max(hp, e);
        diag_src = hup;
        Hprev[c] = h;
        Fprev[c] = f;
        left_Hp = hp;
        left_E = e;
      }
      pub_Hp = left_Hp;
      pub_E = left_E;
      pub_H = Hprev[CPL - 1];
      Hleft_prev = rH;
      bt[(size_t)(i - 1) * 32 + lane] = word;
    }
  }
  int const cm = (m - 1) % CPL;
  int hsel = Hprev[0];
#pragma unroll
  for (int c = 1; c < CPL; ++c)
    if (cm == c)
      hsel = Hprev[c];
  return __shfl_sync(FULL, hsel, last_lane);
}

// Starts the asynchronous copy (LDGSTS) of backtrack rows top, top-1, ... top-SW_ROWS_CACHED+1 of this warp into one
// shared-memory buffer; the traceback consumes one buffer while the next one is in flight.
__device__ __forceinline__ void sw_stage_rows_async(uint32_t (*rows)[32], const uint32_t * bt, int top, int lane)
{
#pragma unroll
  for (int k = 0; k < SW_ROWS_CACHED; ++k)
    if (top - k >= 0)
    {
      unsigned const dst = (unsigned)__cvta_generic_to_shared(&rows[k][lane]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(bt + (size_t)(top - k) * 32 + lane) : "memory");
    }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// The same staging as ONE bulk asynchronous copy per refill (cp.async.bulk, the 1-D form of TMA: SASS UBLKCP): the rows
// base .. top of a warp's slab are contiguous (128 bytes each), so lane 0 arms the buffer's mbarrier with the byte count and
// issues one copy of up to 2 KiB that the copy unit completes on the barrier; the buffer then holds the rows in ASCENDING
// order, row r at index r - (top - SW_ROWS_CACHED + 1).  Returns false when there is no row left to stage.
__device__ __forceinline__ bool sw_stage_rows_bulk(uint32_t (*rows)[32], unsigned long long * bar, const uint32_t * bt, int top, int lane)
{
  if (top < 0)
    return false;
  int const first = top - (SW_ROWS_CACHED - 1);
  int const base = first < 0 ? 0 : first;
  if (lane == 0)
  {
    unsigned const bytes = (unsigned)(top - base + 1) * 128u;
    unsigned const dst = (unsigned)__cvta_generic_to_shared(&rows[base - first][0]);
    unsigned const b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(bt + (size_t)base * 32), "r"(bytes), "r"(b)
                 : "memory");
  }
  return true;
}
__device__ __forceinline__ void sw_wait_rows(unsigned long long * bar, unsigned phase)
{
  unsigned const b = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done;
  do
  {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(b), "r"(phase)
                 : "memory");
  } while (!done);
}

template <bool BULK>
__global__ void __launch_bounds__(SW_WARPS * 32, SW_MIN_BLOCKS) sw_kernel(SwParams P)
{
  __shared__ __align__(128) uint32_t s_rows[SW_WARPS][2][SW_ROWS_CACHED][32];
  __shared__ __align__(8) unsigned long long s_bar[SW_WARPS][2]; // BULK: one mbarrier per staging buffer of every warp
  __shared__ uint8_t s_db[SW_WARPS][GTB_SW_MAX_DATABASE];
  __shared__ uint8_t s_q[SW_WARPS][SW_MAX_Q];
  int const lane = threadIdx.x & 31;
  int const wib = threadIdx.x >> 5;
  int const warp_global = blockIdx.x * SW_WARPS + wib;
  int const total_warps = gridDim.x * SW_WARPS;
  // per-warp scratch slabs are max_db + 5 rows apart: a power-of-two stride would map every warp's row r to the same
  // L2 sets
  uint32_t * bt = P.bt + (size_t)warp_global * (size_t)(P.max_db + 5) * 32;
  unsigned bar_phase[2] = {0u, 0u};
  if (BULK)
  {
    if (lane == 0)
    {
      for (int k = 0; k < 2; ++k)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_bar[wib][k])) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  }

  for (int pair = warp_global; pair < P.n_pairs; pair += total_warps)
  {
    const uint8_t * q = P.q + P.q_off[pair];
    const uint8_t * d = P.d + P.d_off[pair];
    int const m = P.q_off[pair + 1] - P.q_off[pair];
    int const n = P.d_off[pair + 1] - P.d_off[pair];
    if (m <= 0 || m > SW_MAX_Q || n <= 0 || n > P.max_db || n > GTB_SW_MAX_DATABASE)
    {
      if (lane == 0)
        P.out[pair] = gtb_sw_result{0, 0, 0, 0, -1}; // clip_end = -1 marks an unsupported size
      continue;
    }
    // Stage both sequences of this pair in shared memory (the DP reads a database base per step, the traceback
    // compares raw bases).
    __syncwarp();
    for (int k = lane; k < n; k += 32)
      s_db[wib][k] = __ldg(d + k);
    for (int k = lane; k < m; k += 32)
      s_q[wib][k] = __ldg(q + k);
    __syncwarp();
    q = s_q[wib];
    d = s_db[wib];
    int const cpl = (m + 31) / 32;
    int const cpl_inv = (65536 + cpl - 1) / cpl;
    long score;
    switch (cpl)
    {
    case 1: score = sw_dp<1>(q, m, n, d, bt, lane); break;
    case 2: score = sw_dp<2>(q, m, n, d, bt, lane); break;
    case 3: score = sw_dp<3>(q, m, n, d, bt, lane); break;
    case 4: score = sw_dp<4>(q, m, n, d, bt, lane); break;
    default: score = sw_dp<5>(q, m, n, d, bt, lane); break;
    }
    __syncwarp();

    // ---- traceback: clipping (alignment_results.hpp:451-669) and database begin/end (:392-447) follow the same walk.
    // All lanes execute it redundantly on broadcast values; backtrack rows are staged 8 at a time in shared memory.
    long tmp_score = 0, best_begin = 0;
    int res_first = 0, res_second = m, db_first = 0, db_second = n;
    {
      int i = n, j = m;
      // s_rows[wib][cur][k] holds row (cached_top - k); the other buffer is being filled with the SW_ROWS_CACHED rows
      // below.  The walk only ever moves to the same or the previous row.
      int cur = 0, cached_top = n - 1;
      bool pending[2] = {false, false}; // BULK: a copy into the buffer is in flight (its barrier has not been waited for)
      if (BULK)
      {
        // the DP's row stores (generic proxy) precede the copy unit's reads of the slab (async proxy)
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __syncwarp();
        pending[0] = sw_stage_rows_bulk(s_rows[wib][0], &s_bar[wib][0], bt, cached_top, lane);
        pending[1] = sw_stage_rows_bulk(s_rows[wib][1], &s_bar[wib][1], bt, cached_top - SW_ROWS_CACHED, lane);
        sw_wait_rows(&s_bar[wib][0], bar_phase[0]);
        bar_phase[0] ^= 1u;
        pending[0] = false;
      }
      else
      {
        sw_stage_rows_async(s_rows[wib][0], bt, cached_top, lane);
        sw_stage_rows_async(s_rows[wib][1], bt, cached_top - SW_ROWS_CACHED, lane);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
      }
      auto bits_at = [&](int row, int col) -> uint32_t
      {
        if (row <= cached_top - SW_ROWS_CACHED)
        {
          if (BULK)
          {
            int const nxt = cur ^ 1;
            if (pending[nxt])
            {
              sw_wait_rows(&s_bar[wib][nxt], bar_phase[nxt]);
              bar_phase[nxt] ^= 1u;
              pending[nxt] = false;
            }
            __syncwarp(); // every lane is done with the buffer that is refilled next
            cached_top -= SW_ROWS_CACHED;
            pending[cur] = sw_stage_rows_bulk(s_rows[wib][cur], &s_bar[wib][cur], bt, cached_top - SW_ROWS_CACHED, lane);
            cur = nxt;
          }
          else
          {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            cached_top -= SW_ROWS_CACHED;
            sw_stage_rows_async(s_rows[wib][cur], bt, cached_top - SW_ROWS_CACHED, lane);
            cur ^= 1;
          }
        }
        int const wl = ((col - 1) * cpl_inv) >> 16; // (col - 1) / cpl, exact for col <= 160
        // cp.async buffers hold the rows top-down, bulk buffers bottom-up
        uint32_t const w = s_rows[wib][cur][BULK ? SW_ROWS_CACHED - 1 - (cached_top - row) : cached_top - row][wl];
        return (w >> (4 * (cpl - 1 - (col - 1 - wl * cpl)))) & 15u;
      };
      while (i > 0 || j > 0)
      {
        if (j == 0)
        {
          db_first = i;
          break;
        }
        if (i == 0)
        {
          j = 0;
          continue;
        }
        uint32_t const b = bits_at(i - 1, j);
        if (b & 1u) // deletion: query bases against a gap
        {
          while (j > 1 && (bits_at(i - 1, j) & 4u))
          {
            tmp_score -= GE;
            --j;
          }
          tmp_score -= GO;
          --j;
        }
        else if (b & 2u) // insertion: database bases against a gap
        {
          while (i > 1 && (bits_at(i - 1, j) & 8u))
          {
            if (j < m)
              tmp_score -= GE;
            --i;
          }
          --i;
          if (j < m)
            tmp_score -= GO;
          if (j == m)
            db_second = i;
        }
        else
        {
          --i;
          --j;
          if (q[j] == d[i])
          {
            if (tmp_score < 0 - (long)CLIP)
            {
              res_second = j + 1;
              score -= (long)CLIP + tmp_score;
              best_begin += (long)CLIP + tmp_score;
              if (best_begin <= 0)
              {
                best_begin = 0;
                res_first = 0;
              }
              tmp_score = -(long)CLIP;
            }
            tmp_score += MATCH;
            if (tmp_score - (long)CLIP > score)
            {
              long const diff = tmp_score - (long)CLIP - score;
              if (diff > best_begin)
              {
                best_begin = diff;
                res_first = j;
              }
            }
          }
          else
            tmp_score -= MISMATCH;
        }
      }
      if (BULK) // copies the walk did not need any more: their barriers complete a phase all the same
        for (int k = 0; k < 2; ++k)
          if (pending[k])
          {
            sw_wait_rows(&s_bar[wib][k], bar_phase[k]);
            bar_phase[k] ^= 1u;
          }
    }
    if (!BULK)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0)
    {
      gtb_sw_result r;
      r.score = (int32_t)(tmp_score + best_begin);
      r.database_begin = db_first;
      r.database_end = db_second;
      r.clip_begin = res_first;
      r.clip_end = res_second;
      P.out[pair] = r;
    }
    __syncwarp();
  }
}

// Resident warps for windows of up to max_db bases.  Every resident warp owns a backtrack slab of (max_db + 5) rows
// x 128 B that is written once by the DP and read back by the traceback; measured on B200 the kernel is fastest when
// all slabs together stay inside the L2 (12 warps / SM at 512-base windows, 116 MB of 126 MB) -- more warps spill the
// slabs to HBM and lose more to the traceback's read latency than they gain in issue slots.
int sw_resident_warps(int max_db)
{
  int dev = 0, sms = 148, nb = 0, l2 = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sw_kernel<false>, SW_WARPS * 32, 0) != cudaSuccess || nb < 1)
    nb = 1;
  if (nb > SW_MIN_BLOCKS)
    nb = SW_MIN_BLOCKS;
  if (l2 > 0)
  {
    size_t const slab = (size_t)(max_db + 5) * 128;
    int const fit = (int)((size_t)l2 * 95 / 100 / slab / (size_t)(sms * SW_WARPS));
    nb = max(1, min(nb, fit));
  }
  if (const char * e = getenv("GTB_SW_BLOCKS")) // tuning knob: resident blocks per SM
    nb = max(1, min(SW_MIN_BLOCKS, atoi(e)));
  return sms * nb * SW_WARPS;
}

void launch_sw(const SwParams & p, int resident_warps, void * stream)
{
  if (p.n_pairs <= 0)
    return;
  int grid = resident_warps / SW_WARPS;
  int const need = (p.n_pairs + SW_WARPS - 1) / SW_WARPS;
  if (need < grid)
    grid = need;
  // Traceback rows are staged by bulk asynchronous copies (cp.async.bulk + mbarrier): measured 13.57 -> 12.96 ms for 1e5 pairs
  // (434 -> 455 GCUPS), bit-identical on the parity sets.  GTB_SW_BULK=0 selects the per-lane cp.async staging.
  static bool const bulk = []() { const char * e = getenv("GTB_SW_BULK"); return !e || atoi(e) != 0; }();
  if (bulk)
    sw_kernel<true><<<grid, SW_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  else
    sw_kernel<false><<<grid, SW_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
}

} // namespace gtb
