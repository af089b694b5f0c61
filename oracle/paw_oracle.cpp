// oracle/paw_oracle.cpp -- TEST INFRASTRUCTURE ONLY: scalar restatement of paw::pairwise_alignment as graphtyper's
// discovery re-alignment configures it (src/typer/caller.cpp:1864-1870): semi-global (query global, both database ends
// free), affine gaps (open 7 = cost of the first gap base, extend 1), match +1 / mismatch -4, then soft clipping
// (penalty 5) applied along the traceback.  Follows paw/include/paw/align/pairwise_alignment.hpp:146-388 (recurrences
// and the strict-greater tie rules that define the backtrack bits), alignment_results.hpp:392-447 (database begin/end)
// and :451-669 (clipping).  Pinned against the compiled paw (oracle/_ref/bin/paw_probe) by tests/test_sw_oracle.py.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

struct gto_sw_result
{
  int32_t score, database_begin, database_end, clip_begin, clip_end;
};

static inline bool is_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
static inline char db_upper(char c) // magic_function (libsimdpp_utils.hpp:96-123) is case-insensitive for the database
{
  switch (c)
  {
  case 'a': return 'A';
  case 'c': return 'C';
  case 'g': return 'G';
  case 't': return 'T';
  default: return c;
  }
}

int gto_sw_align(const char * q, int m, const char * d, int n, gto_sw_result * out)
{
  int const MATCH = 1, MISMATCH = 4, GO = 7, GE = 1, CLIP = 5;
  int const NEG = -(1 << 28);
  // bits per cell: 1 del, 2 ins, 4 del_extend, 8 ins_extend   (libsimdpp_backtracker.hpp:33-41)
  std::vector<uint8_t> bt((size_t)n * (m + 1), 0);
  std::vector<int> H(m + 1), F(m + 1, NEG), Hn(m + 1), Fn(m + 1), Hp(m + 1);
  H[0] = 0;
  for (int j = 1; j <= m; ++j)
    H[j] = -(GO + (j - 1) * GE); // initial row: alignment_options.hpp:282-299 in the un-gained domain
  for (int i = 1; i <= n; ++i)
  {
    uint8_t * b = bt.data() + (size_t)(i - 1) * (m + 1);
    char const dc = db_upper(d[i - 1]);
    // column 0: free leading database bases (left_column_free)
    Fn[0] = H[0];
    Hp[0] = H[0];
    Hn[0] = H[0];
    int E = NEG;
    for (int j = 1; j <= m; ++j)
    {
      bool const eq = is_acgt(dc) && q[j - 1] == dc; // W profile: alignment_cache.hpp:70-121
      int const diag = H[j - 1] + (eq ? MATCH : -MISMATCH);
      int fopen = H[j] - GO;
      if (j == m)
        fopen = H[j]; // right_column_free: trailing database bases cost nothing
      int const fext = F[j] - GE;
      uint8_t bits = 0;
      int f = fopen;
      if (fext > fopen)
      {
        f = fext;
        bits |= 8;
      }
      int hp = diag;
      if (f > diag)
      {
        hp = f;
        bits |= 2;
      }
      int const eopen = Hp[j - 1] - GO;
      int const eext = E - GE;
      int e = eopen;
      if (eext > eopen)
      {
        e = eext;
        bits |= 4;
      }
      E = e;
      int h = hp;
      if (e > hp)
      {
        h = e;
        bits |= 1;
      }
      Hp[j] = hp;
      Hn[j] = h;
      Fn[j] = f;
      b[j] = bits;
    }
    H.swap(Hn);
    F.swap(Fn);
  }
  long score = H[m];
  auto B = [&](long i, long j) { return bt[(size_t)i * (m + 1) + j]; };

  // ---- apply_clipping (alignment_results.hpp:451-669), clip_left = clip_right = CLIP
  long best_begin_clip_improvement = 0, tmp_score = 0;
  long res_first = 0, res_second = m;
  {
    long i = n, j = m;
    while (i > 0 || j > 0)
    {
      if (j == 0)
        break;
      if (i == 0)
      {
        j = 0;
      }
      else if (B(i - 1, j) & 1)
      {
        while (j > 1 && (B(i - 1, j) & 4))
        {
          tmp_score -= GE;
          --j;
        }
        tmp_score -= GO;
        --j;
      }
      else if (B(i - 1, j) & 2)
      {
        while (i > 1 && (B(i - 1, j) & 8))
        {
          if (j < m)
            tmp_score -= GE;
          --i;
        }
        --i;
        if (j < m)
          tmp_score -= GO;
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
            best_begin_clip_improvement += (long)CLIP + tmp_score;
            if (best_begin_clip_improvement <= 0)
            {
              best_begin_clip_improvement = 0;
              res_first = 0;
            }
            tmp_score = -(long)CLIP;
          }
          tmp_score += MATCH;
          if (tmp_score - (long)CLIP > score)
          {
            long const diff = tmp_score - (long)CLIP - score;
            if (diff > best_begin_clip_improvement)
            {
              best_begin_clip_improvement = diff;
              res_first = j;
            }
          }
        }
        else
          tmp_score -= MISMATCH;
      }
    }
  }
  score = tmp_score + best_begin_clip_improvement;

  // ---- get_database_begin_end (alignment_results.hpp:392-447)
  long db_first = 0, db_second = n;
  {
    long i = n, j = m;
    while (i > 0 || j > 0)
    {
      if (j == 0)
      {
        db_first = i;
        break;
      }
      if (i == 0)
        j = 0;
      else if (B(i - 1, j) & 1)
      {
        while (j > 1 && (B(i - 1, j) & 4))
          --j;
        --j;
      }
      else if (B(i - 1, j) & 2)
      {
        while (i > 1 && (B(i - 1, j) & 8))
          --i;
        --i;
        if (j == m)
          db_second = i;
      }
      else
      {
        --i;
        --j;
      }
    }
  }
  out->score = (int32_t)score;
  out->database_begin = (int32_t)db_first;
  out->database_end = (int32_t)db_second;
  out->clip_begin = (int32_t)res_first;
  out->clip_end = (int32_t)res_second;
  return 0;
}

int gto_sw_align_batch(int n_pairs, const char * qs, const int32_t * q_off, const char * ds, const int32_t * d_off,
                       gto_sw_result * out)
{
  for (int k = 0; k < n_pairs; ++k)
    gto_sw_align(qs + q_off[k], q_off[k + 1] - q_off[k], ds + d_off[k], d_off[k + 1] - d_off[k], out + k);
  return 0;
}
}
