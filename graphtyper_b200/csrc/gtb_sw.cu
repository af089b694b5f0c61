// gtb_sw.cu -- discovery re-alignment kernel: B200 replacement of paw::pairwise_alignment as
// graphtyper's realign_to_indels calls it (src/typer/caller.cpp:1864-1870,2007; paw/include/paw/align/
// pairwise_alignment.hpp:146-388): semi-global affine-gap alignment of a read (query, global) against a haplotype
// window (database, both ends free), match +1 / mismatch -4 / gap open 7 (first gap base) / extend 1, followed by the
// soft-clipping pass (penalty 5) and database begin/end extraction along the traceback
// (alignment_results.hpp:392-447,451-669).  Results are bit-identical to paw: score, database_begin/end, clip_begin/end.
//
// One warp per (read, window) pair.  The DP runs as a 32-lane systolic wavefront: lane l owns query columns
// 5l+1 .. 5l+5, at step s it computes database row s-l+1, and hands (H', E, H) of its last column to lane l+1 with one
// shuffle per step.  Every cell is three fused max-add DPX operations (__viaddmax_s32 / __vibmax_s32) plus the
// strict-greater predicates that define paw's four backtrack bits (del, ins, del_extend, ins_extend); the bits of a
// lane's five cells are one 32-bit word per row (coalesced 128-byte row stores to a per-warp scratch).  The traceback
// (clipping + begin/end) is executed redundantly by all lanes on rows staged eight at a time in shared memory.
// No tensor cores: this is min/max/add dynamic programming, not a contraction.

#include <cstdint>
#include <cuda_runtime.h>

#include "gtb_device.cuh"

namespace gtb
{
constexpr int SW_CPL = 5;                 // query columns per lane
constexpr int SW_MAX_Q = 32 * SW_CPL;     // 160 >= MAX_READ_LENGTH (151)
constexpr int SW_WARPS = 4;               // warps per block
constexpr int SW_ROWS_CACHED = 8;         // backtrack rows staged per refill
constexpr int SW_NEG = -(1 << 28);

namespace
{
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int MATCH = 1, MISMATCH = 4, GO = 7, GE = 1, CLIP = 5;

__device__ __forceinline__ uint8_t db_upper(uint8_t c) // magic_function is case-insensitive (libsimdpp_utils.hpp:96-123)
{
  return (c == 'a' || c == 'c' || c == 'g' || c == 't') ? (uint8_t)(c - 32) : c;
}
__device__ __forceinline__ bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
} // namespace

__global__ void __launch_bounds__(SW_WARPS * 32) sw_kernel(SwParams P)
{
  __shared__ uint32_t s_rows[SW_WARPS][SW_ROWS_CACHED][32];
  int const lane = threadIdx.x & 31;
  int const wib = threadIdx.x >> 5;
  int const warp_global = blockIdx.x * SW_WARPS + wib;
  int const total_warps = gridDim.x * SW_WARPS;
  uint32_t * bt = P.bt + (size_t)warp_global * (size_t)P.max_db * 32;

  for (int pair = warp_global; pair < P.n_pairs; pair += total_warps)
  {
    const uint8_t * q = P.q + P.q_off[pair];
    const uint8_t * d = P.d + P.d_off[pair];
    int const m = P.q_off[pair + 1] - P.q_off[pair];
    int const n = P.d_off[pair + 1] - P.d_off[pair];
    if (m <= 0 || m > SW_MAX_Q || n <= 0 || n > P.max_db)
    {
      if (lane == 0)
        P.out[pair] = gtb_sw_result{0, 0, 0, 0, -1}; // clip_end = -1 marks an unsupported size
      continue;
    }
    int const c0 = lane * SW_CPL; // columns c0+1 .. c0+5
    uint8_t qc[SW_CPL];
    int Hprev[SW_CPL], Fprev[SW_CPL];
#pragma unroll
    for (int c = 0; c < SW_CPL; ++c)
    {
      int const j = c0 + c + 1;
      qc[c] = j <= m ? q[j - 1] : (uint8_t)0;
      Hprev[c] = -(GO + (j - 1) * GE); // initial row (alignment_options.hpp:282-299)
      Fprev[c] = SW_NEG;
    }
    int Hleft_prev = lane == 0 ? 0 : -(GO + (c0 - 1) * GE); // H[0][c0]
    int pub_Hp = 0, pub_E = SW_NEG, pub_H = 0;
    int const steps = n + 31;
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
      if (i >= 1 && i <= n && c0 < m)
      {
        uint8_t const dc = db_upper(__ldg(d + i - 1));
        bool const dvalid = is_acgt(dc);
        int left_Hp = rHp, left_E = rE, diag_src = Hleft_prev;
        uint32_t word = 0;
#pragma unroll
        for (int c = 0; c < SW_CPL; ++c)
        {
          int const j = c0 + c + 1;
          if (j <= m)
          {
            int const diag = diag_src + ((dvalid && qc[c] == dc) ? MATCH : -MISMATCH);
            int const hup = Hprev[c];
            int const fopen = (j == m) ? hup : hup - GO; // right_column_free
            bool p;
            // F = max(F_up - extend, open); ins_extend bit <=> extension strictly better
            int const f = __vibmax_s32(fopen, Fprev[c] - GE, &p);
            uint32_t bits = p ? 0u : 8u;
            int const hp = __vibmax_s32(diag, f, &p);
            bits |= p ? 0u : 2u;
            int const e = __vibmax_s32(left_Hp - GO, left_E - GE, &p);
            bits |= p ? 0u : 4u;
            int const h = __vibmax_s32(hp, e, &p);
            bits |= p ? 0u : 1u;
            diag_src = hup;
            Hprev[c] = h;
            Fprev[c] = f;
            left_Hp = hp;
            left_E = e;
            word |= bits << (4 * c);
          }
        }
        pub_Hp = left_Hp;
        pub_E = left_E;
        // H of this lane's last column (only lanes whose last column is <= m feed a neighbour that is active)
        pub_H = Hprev[SW_CPL - 1];
        Hleft_prev = rH;
        bt[(size_t)(i - 1) * 32 + lane] = word;
      }
    }
    // DP score = H[n][m]
    int const lm = (m - 1) / SW_CPL, cm = (m - 1) % SW_CPL;
    int hsel = Hprev[0];
#pragma unroll
    for (int c = 1; c < SW_CPL; ++c)
      if (cm == c)
        hsel = Hprev[c];
    long score = __shfl_sync(FULL, hsel, lm);
    __syncwarp();

    // ---- traceback: clipping (alignment_results.hpp:451-669) and database begin/end (:392-447) follow the same walk.
    // All lanes execute it redundantly on broadcast values; backtrack rows are staged 8 at a time in shared memory.
    long tmp_score = 0, best_begin = 0;
    int res_first = 0, res_second = m, db_first = 0, db_second = n;
    {
      int i = n, j = m;
      int cached_top = -1; // s_rows[k] holds row (cached_top - k)
      auto bits_at = [&](int row, int col) -> uint32_t
      {
        if (cached_top < 0 || row > cached_top || row <= cached_top - SW_ROWS_CACHED)
        {
          __syncwarp();
#pragma unroll
          for (int k = 0; k < SW_ROWS_CACHED; ++k)
            s_rows[wib][k][lane] = (row - k) >= 0 ? bt[(size_t)(row - k) * 32 + lane] : 0u;
          cached_top = row;
          __syncwarp();
        }
        uint32_t const w = s_rows[wib][cached_top - row][(col - 1) / SW_CPL];
        return (w >> (4 * ((col - 1) % SW_CPL))) & 15u;
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
    }
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

int sw_resident_warps()
{
  int dev = 0, sms = 148, nb = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sw_kernel, SW_WARPS * 32, 0) != cudaSuccess || nb < 1)
    nb = 1;
  if (nb > 4)
    nb = 4; // 16 warps / SM: bounds the backtrack scratch (128 B per row per warp)
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
  sw_kernel<<<grid, SW_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
}

} // namespace gtb
