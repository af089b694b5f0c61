// gtb_inflate.cuh -- DEFLATE (RFC 1951) decoder + CRC-32 for BGZF blocks, one WARP per block (product code; SURVEY.md
// section 8f, N3: "BGZF inflate" of src/utilities/hts_reader.cpp:166-303 -> htslib bgzf.c / libdeflate on the reference side).
//
// The same source compiles for the host, where a "warp" is one lane: tests/test_bgzf_host.py runs exactly this code on the CPU
// (gtb_debug_bgzf_host) against zlib, so what the GPU run adds is only the 32-lane cooperation:
//   * lane 0 owns the bit reader and decodes symbols (canonical Huffman: a first-level table of 2^10 / 2^8 entries indexed by
//     the next bits of the LSB-first stream, longer codes by the bit-serial canonical walk), writes literals itself,
//   * every match (length, distance) and every stored block is broadcast and copied by all lanes (a distance shorter than the
//     length repeats with period `distance`, so lane i reads out[pos - distance + i % distance]: no lane depends on a byte
//     written by the same copy),
//   * CRC-32: every lane takes one 32nd of the inflated bytes, the partial CRCs are folded with the x^(8n) mod P shift
//     (the crc32_combine identity), so the check htslib makes on every block (bgzf.c: inflate_block / check) is kept.
// Tables live in shared memory on the device (3.2 KiB per warp).
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define GTB_HD __host__ __device__ __forceinline__
#define GTB_HDN inline __host__ __device__ __noinline__
#else
#define GTB_HD inline
#define GTB_HDN inline
#endif

namespace gtb
{
constexpr int INF_LIT_FAST = 10, INF_DIST_FAST = 8;
constexpr int INF_OK = 0, INF_ERR_INPUT = -1 /* ran out of input */, INF_ERR_TYPE = -2 /* block type 3 */,
              INF_ERR_STORED = -3 /* LEN != ~NLEN */, INF_ERR_CODE = -4 /* bad code lengths / over-subscribed / incomplete */,
              INF_ERR_SYMBOL = -5 /* invalid symbol in the stream */, INF_ERR_DIST = -6 /* distance before the start */,
              INF_ERR_OUTPUT = -7 /* more output than the block header announced */, INF_ERR_SIZE = -8 /* ISIZE mismatch */,
              INF_ERR_CRC = -9, INF_ERR_HEADER = -10 /* not a BGZF block header */;

struct InflateTables
{
  uint16_t lit_fast[1 << INF_LIT_FAST];   // (symbol << 4) | code length, 0 = the code is longer than the table's bits
  uint16_t dist_fast[1 << INF_DIST_FAST];
  uint16_t lit_count[16], dist_count[16]; // number of codes of every length
  uint16_t lit_sym[288], dist_sym[32];    // symbols in canonical order
};

struct BitReader
{
  const uint8_t * ip;
  const uint8_t * ie;
  unsigned long long buf;
  int cnt;
  int overrun; // bits were requested beyond the end of the input
};

GTB_HD void br_refill(BitReader & b)
{
  // four bytes at once while they are there (four independent loads, one latency), single bytes at the end of the input
  if (b.cnt <= 32 && b.ie - b.ip >= 4)
  {
    uint32_t const w = (uint32_t)b.ip[0] | ((uint32_t)b.ip[1] << 8) | ((uint32_t)b.ip[2] << 16) | ((uint32_t)b.ip[3] << 24);
    b.buf |= (unsigned long long)w << b.cnt;
    b.cnt += 32;
    b.ip += 4;
    return;
  }
  while (b.cnt <= 56 && b.ip < b.ie)
  {
    b.buf |= (unsigned long long)(*b.ip++) << b.cnt;
    b.cnt += 8;
  }
}
// the next n (<= 32) bits, LSB first; zero bits beyond the end of the input (flagged)
GTB_HD uint32_t br_bits(BitReader & b, int n)
{
  if (b.cnt < n)
  {
    br_refill(b);
    if (b.cnt < n)
    {
      b.overrun = 1;
      b.cnt = n; // zeros
    }
  }
  uint32_t const v = (uint32_t)(b.buf & ((1ull << n) - 1ull));
  b.buf >>= n;
  b.cnt -= n;
  return v;
}

GTB_HD uint32_t reverse_bits(uint32_t v, int n)
{
  uint32_t r = 0;
  for (int i = 0; i < n; ++i)
  {
    r = (r << 1) | (v & 1u);
    v >>= 1;
  }
  return r;
}

// Canonical Huffman tables from code lengths (the construction of zlib's puff.c, plus the first-level lookup table).
// Returns 0 for a complete code, > 0 for an incomplete one (the caller decides whether that is allowed), < 0 over-subscribed.
GTB_HDN int build_huffman(const uint8_t * length, int n, uint16_t * count, uint16_t * symbol, uint16_t * fast, int fast_bits)
{
  for (int len = 0; len <= 15; ++len)
    count[len] = 0;
  for (int s = 0; s < n; ++s)
    ++count[length[s]];
  for (int i = 0; i < (1 << fast_bits); ++i)
    fast[i] = 0;
  if (count[0] == n)
    return 0; // no codes: complete, but decoding anything fails
  int left = 1;
  for (int len = 1; len <= 15; ++len)
  {
    left <<= 1;
    left -= count[len];
    if (left < 0)
      return left;
  }
  uint16_t offs[16];
  offs[1] = 0;
  for (int len = 1; len < 15; ++len)
    offs[len + 1] = (uint16_t)(offs[len] + count[len]);
  for (int s = 0; s < n; ++s)
    if (length[s] != 0)
      symbol[offs[length[s]]++] = (uint16_t)s;
  // first-level table: canonical codes in order of (length, symbol); the stream carries them bit-reversed
  uint32_t code = 0;
  int idx = 0;
  for (int len = 1; len <= 15; ++len)
  {
    for (int k = 0; k < count[len]; ++k, ++idx, ++code)
      if (len <= fast_bits)
      {
        uint32_t const rev = reverse_bits(code, len);
        uint16_t const entry = (uint16_t)((symbol[idx] << 4) | len);
        for (uint32_t fill = rev; fill < (1u << fast_bits); fill += (1u << len))
          fast[fill] = entry;
      }
    code <<= 1;
  }
  return left;
}

// One symbol: first-level table, then the bit-serial canonical walk for codes beyond it.  < 0: no such code.
GTB_HD int decode_symbol(BitReader & b, const uint16_t * count, const uint16_t * symbol, const uint16_t * fast, int fast_bits)
{
  if (b.cnt < 15)
    br_refill(b);
  uint16_t const e = fast[(uint32_t)b.buf & ((1u << fast_bits) - 1u)];
  if (e != 0)
  {
    int const len = e & 15;
    if (len > b.cnt)
    {
      b.overrun = 1;
      return -1;
    }
    b.buf >>= len;
    b.cnt -= len;
    return e >> 4;
  }
  int code = 0, first = 0, index = 0;
  unsigned long long bits = b.buf;
  for (int len = 1; len <= 15; ++len)
  {
    if (len > b.cnt)
    {
      b.overrun = 1;
      return -1;
    }
    code |= (int)(bits & 1ull);
    bits >>= 1;
    int const cnt = count[len];
    if (code - cnt < first)
    {
      b.buf >>= len;
      b.cnt -= len;
      return symbol[index + (code - first)];
    }
    index += cnt;
    first += cnt;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

// lanes of the cooperating group
#if defined(__CUDA_ARCH__)
#define GTB_INF_LANE (threadIdx.x & 31u)
#define GTB_INF_WIDTH 32u
#define GTB_INF_BCAST(x) __shfl_sync(0xFFFFFFFFu, (x), 0)
#define GTB_INF_SYNC() __syncwarp()
#else
#define GTB_INF_LANE 0u
#define GTB_INF_WIDTH 1u
#define GTB_INF_BCAST(x) (x)
#define GTB_INF_SYNC() ((void)0)
#endif

// Raw DEFLATE stream -> out[0 .. out_cap).  All lanes of the warp call it with the same arguments; *out_len and the return
// value are the same on every lane.
GTB_HDN int inflate_raw(const uint8_t * in, uint32_t in_len, uint8_t * out, uint32_t out_cap, InflateTables & T, uint32_t * out_len)
{
  constexpr uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  constexpr uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  constexpr uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  constexpr uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  constexpr uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  unsigned const lane = GTB_INF_LANE;
  BitReader b{in, in + in_len, 0ull, 0, 0};
  uint32_t op = 0;
  int err = INF_OK;
  int last = 0;
  while (!last && err == INF_OK)
  {
    int type = 0;
    uint32_t stored_len = 0, stored_src = 0;
    if (lane == 0)
    {
      last = (int)br_bits(b, 1);
      type = (int)br_bits(b, 2);
      if (b.overrun)
        err = INF_ERR_INPUT;
      else if (type == 3)
        err = INF_ERR_TYPE;
      else if (type == 0)
      {
        // stored: drop the rest of the current byte, give whole prefetched bytes back
        int const drop = b.cnt & 7;
        b.buf >>= drop;
        b.cnt -= drop;
        b.ip -= b.cnt / 8;
        b.buf = 0;
        b.cnt = 0;
        if (b.ie - b.ip < 4)
          err = INF_ERR_INPUT;
        else
        {
          uint32_t const len = (uint32_t)b.ip[0] | ((uint32_t)b.ip[1] << 8), nlen = (uint32_t)b.ip[2] | ((uint32_t)b.ip[3] << 8);
          b.ip += 4;
          if (len != (~nlen & 0xFFFFu))
            err = INF_ERR_STORED;
          else if ((uint32_t)(b.ie - b.ip) < len)
            err = INF_ERR_INPUT;
          else if (op + len > out_cap)
            err = INF_ERR_OUTPUT;
          else
          {
            stored_len = len;
            stored_src = (uint32_t)(b.ip - in);
            b.ip += len;
          }
        }
      }
      else if (type == 1)
      {
        uint8_t lengths[288];
        for (int s = 0; s < 144; ++s)
          lengths[s] = 8;
        for (int s = 144; s < 256; ++s)
          lengths[s] = 9;
        for (int s = 256; s < 280; ++s)
          lengths[s] = 7;
        for (int s = 280; s < 288; ++s)
          lengths[s] = 8;
        build_huffman(lengths, 288, T.lit_count, T.lit_sym, T.lit_fast, INF_LIT_FAST);
        for (int s = 0; s < 30; ++s)
          lengths[s] = 5;
        build_huffman(lengths, 30, T.dist_count, T.dist_sym, T.dist_fast, INF_DIST_FAST);
      }
      else
      {
        uint8_t lengths[320];
        int const nlen = (int)br_bits(b, 5) + 257, ndist = (int)br_bits(b, 5) + 1, ncode = (int)br_bits(b, 4) + 4;
        if (nlen > 286 || ndist > 30)
          err = INF_ERR_CODE;
        else
        {
          for (int i = 0; i < 19; ++i)
            lengths[CL_ORDER[i]] = i < ncode ? (uint8_t)br_bits(b, 3) : (uint8_t)0;
          // the code-length code decodes through the literal tables' storage (rebuilt right after)
          if (build_huffman(lengths, 19, T.lit_count, T.lit_sym, T.lit_fast, INF_LIT_FAST) != 0)
            err = INF_ERR_CODE;
          int index = 0;
          while (err == INF_OK && index < nlen + ndist)
          {
            int const sym = decode_symbol(b, T.lit_count, T.lit_sym, T.lit_fast, INF_LIT_FAST);
            if (sym < 0)
              err = b.overrun ? INF_ERR_INPUT : INF_ERR_SYMBOL;
            else if (sym < 16)
              lengths[index++] = (uint8_t)sym;
            else
            {
              int len = 0, rep;
              if (sym == 16)
              {
                if (index == 0)
                {
                  err = INF_ERR_CODE;
                  break;
                }
                len = lengths[index - 1];
                rep = 3 + (int)br_bits(b, 2);
              }
              else if (sym == 17)
                rep = 3 + (int)br_bits(b, 3);
              else
                rep = 11 + (int)br_bits(b, 7);
              if (index + rep > nlen + ndist)
              {
                err = INF_ERR_CODE;
                break;
              }
              while (rep--)
                lengths[index++] = (uint8_t)len;
            }
          }
          if (err == INF_OK && b.overrun)
            err = INF_ERR_INPUT;
          if (err == INF_OK && lengths[256] == 0)
            err = INF_ERR_CODE; // no end-of-block code
          if (err == INF_OK)
          {
            uint8_t dl[32];
            for (int i = 0; i < ndist; ++i)
              dl[i] = lengths[nlen + i];
            int e1 = build_huffman(lengths, nlen, T.lit_count, T.lit_sym, T.lit_fast, INF_LIT_FAST);
            if (e1 != 0 && (e1 < 0 || nlen != T.lit_count[0] + T.lit_count[1]))
              err = INF_ERR_CODE; // incomplete codes are allowed only when there is a single one-bit code
            int e2 = build_huffman(dl, ndist, T.dist_count, T.dist_sym, T.dist_fast, INF_DIST_FAST);
            if (e2 != 0 && (e2 < 0 || ndist != T.dist_count[0] + T.dist_count[1]))
              err = INF_ERR_CODE;
          }
        }
      }
    }
    err = GTB_INF_BCAST(err);
    last = GTB_INF_BCAST(last);
    type = GTB_INF_BCAST(type);
    if (err != INF_OK)
      break;
    if (type == 0)
    {
      stored_len = GTB_INF_BCAST(stored_len);
      stored_src = GTB_INF_BCAST(stored_src);
      for (uint32_t i = lane; i < stored_len; i += GTB_INF_WIDTH)
        out[op + i] = in[stored_src + i];
      op += stored_len;
      GTB_INF_SYNC();
      continue;
    }
    // ---- compressed block: lane 0 decodes up to the next match; matches are copied by all lanes
    for (;;)
    {
      uint32_t kind = 0, mlen = 0, mdist = 0, at = op; // kind: 0 match, 1 end of block, 2 error
      if (lane == 0)
      {
        for (;;)
        {
          int const sym = decode_symbol(b, T.lit_count, T.lit_sym, T.lit_fast, INF_LIT_FAST);
          if (sym < 0)
          {
            err = b.overrun ? INF_ERR_INPUT : INF_ERR_SYMBOL;
            kind = 2;
            break;
          }
          if (sym < 256)
          {
            if (at >= out_cap)
            {
              err = INF_ERR_OUTPUT;
              kind = 2;
              break;
            }
            out[at++] = (uint8_t)sym;
            continue;
          }
          if (sym == 256)
          {
            kind = 1;
            break;
          }
          int const ls = sym - 257;
          if (ls >= 29)
          {
            err = INF_ERR_SYMBOL;
            kind = 2;
            break;
          }
          mlen = LEN_BASE[ls] + br_bits(b, LEN_EXTRA[ls]);
          int const ds = decode_symbol(b, T.dist_count, T.dist_sym, T.dist_fast, INF_DIST_FAST);
          if (ds < 0 || ds >= 30)
          {
            err = b.overrun ? INF_ERR_INPUT : INF_ERR_SYMBOL;
            kind = 2;
            break;
          }
          mdist = DIST_BASE[ds] + br_bits(b, DIST_EXTRA[ds]);
          if (b.overrun)
          {
            err = INF_ERR_INPUT;
            kind = 2;
          }
          else if (mdist > at)
          {
            err = INF_ERR_DIST;
            kind = 2;
          }
          else if (at + mlen > out_cap)
          {
            err = INF_ERR_OUTPUT;
            kind = 2;
          }
          break;
        }
      }
      kind = GTB_INF_BCAST(kind);
      at = GTB_INF_BCAST(at);
      op = at;
      if (kind != 0)
        break;
      mlen = GTB_INF_BCAST(mlen);
      mdist = GTB_INF_BCAST(mdist);
      GTB_INF_SYNC(); // lane 0's literals are visible to the lanes that copy
      uint32_t const from = op - mdist;
      if (mdist >= mlen)
        for (uint32_t i = lane; i < mlen; i += GTB_INF_WIDTH)
          out[op + i] = out[from + i];
      else // the match overlaps its own output: period mdist
        for (uint32_t i = lane; i < mlen; i += GTB_INF_WIDTH)
          out[op + i] = out[from + i % mdist];
      op += mlen;
      GTB_INF_SYNC();
    }
    err = GTB_INF_BCAST(err);
  }
  *out_len = op;
  return err;
}

// ---- CRC-32 (IEEE 802.3, reflected, as gzip)
GTB_HD uint32_t crc32_update(uint32_t crc, const uint8_t * p, uint32_t n)
{
  constexpr uint32_t NIB[16] = {0x00000000u, 0x1DB71064u, 0x3B6E20C8u, 0x26D930ACu, 0x76DC4190u, 0x6B6B51F4u, 0x4DB26158u, 0x5005713Cu,
                                0xEDB88320u, 0xF00F9344u, 0xD6D6A3E8u, 0xCB61B38Cu, 0x9B64C2B0u, 0x86D3D2D4u, 0xA00AE278u, 0xBDBDF21Cu};
  crc = ~crc;
  uint32_t i = 0;
  for (; i < n && (reinterpret_cast<uintptr_t>(p + i) & 3u) != 0; ++i)
  {
    crc ^= p[i];
    crc = NIB[crc & 15u] ^ (crc >> 4);
    crc = NIB[crc & 15u] ^ (crc >> 4);
  }
  // aligned body, a little-endian word at a time: one load per four bytes, and the loads do not depend on the CRC chain
  for (; i + 4 <= n; i += 4)
  {
    crc ^= *reinterpret_cast<const uint32_t *>(p + i);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      crc = NIB[crc & 15u] ^ (crc >> 4);
  }
  for (; i < n; ++i)
  {
    crc ^= p[i];
    crc = NIB[crc & 15u] ^ (crc >> 4);
    crc = NIB[crc & 15u] ^ (crc >> 4);
  }
  return ~crc;
}
// a(x) * b(x) mod P(x), reflected representation (zlib crc32.c: multmodp)
GTB_HD uint32_t crc32_multmodp(uint32_t a, uint32_t b)
{
  uint32_t m = 1u << 31, p = 0;
  for (;;)
  {
    if (a & m)
    {
      p ^= b;
      if ((a & (m - 1u)) == 0)
        break;
    }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
  }
  return p;
}
// x^(8 * n) mod P(x)
GTB_HD uint32_t crc32_shift_of(uint32_t n_bytes)
{
  uint32_t sq = 0x40000000u; // x^1
  for (int i = 0; i < 3; ++i)
    sq = crc32_multmodp(sq, sq); // x^8
  uint32_t p = 1u << 31;         // x^0
  while (n_bytes)
  {
    if (n_bytes & 1u)
      p = crc32_multmodp(sq, p);
    sq = crc32_multmodp(sq, sq);
    n_bytes >>= 1;
  }
  return p;
}
// crc of A||B from crc(A), crc(B) and shift = crc32_shift_of(len(B))
GTB_HD uint32_t crc32_combine_shift(uint32_t crc_a, uint32_t crc_b, uint32_t shift) { return crc32_multmodp(shift, crc_a) ^ crc_b; }

// CRC-32 of out[0 .. n) by the cooperating lanes; the same value on every lane.
GTB_HD uint32_t crc32_coop(const uint8_t * p, uint32_t n)
{
#if defined(__CUDA_ARCH__)
  // crc(A|B|C) = shift(crc A, |B| + |C|) ^ shift(crc B, |C|) ^ crc C (the shift is linear): every lane shifts the CRC of its
  // own piece by the bytes that follow it, the warp XORs the results
  unsigned const lane = threadIdx.x & 31u;
  uint32_t const per = (n + 31u) / 32u;
  uint32_t const lo = lane * per < n ? lane * per : n, hi = lo + per < n ? lo + per : n;
  uint32_t mine = hi > lo ? crc32_update(0u, p + lo, hi - lo) : 0u;
  if (hi > lo && hi < n)
    mine = crc32_multmodp(crc32_shift_of(n - hi), mine);
  for (int d = 16; d >= 1; d >>= 1)
    mine ^= __shfl_xor_sync(0xFFFFFFFFu, mine, d);
  return mine;
#else
  // the host build folds in pieces as well, so that the combine identity is what the CPU tests exercise
  uint32_t const per = (n + 31u) / 32u;
  uint32_t crc = 0;
  for (uint32_t lo = 0; lo < n; lo += per)
  {
    uint32_t const len = lo + per < n ? per : n - lo;
    crc = crc32_combine_shift(crc, crc32_update(0u, p + lo, len), crc32_shift_of(len));
  }
  return crc;
#endif
}

// ---- one BGZF block (RFC 1952 member with the BC extra field, SAM spec 4.1): header checks, inflate, ISIZE, CRC-32
struct BgzfBlockInfo
{
  uint32_t header_bytes; // offset of the DEFLATE stream
  uint32_t block_bytes;  // BSIZE + 1
  uint32_t isize;
  uint32_t crc;
};
// Parses the header of the block starting at p (avail bytes readable).  0 or INF_ERR_HEADER / INF_ERR_INPUT.
GTB_HD int bgzf_block_info(const uint8_t * p, unsigned long long avail, BgzfBlockInfo * info)
{
  if (avail < 18)
    return INF_ERR_INPUT;
  if (p[0] != 31 || p[1] != 139 || p[2] != 8 || (p[3] & 4) == 0)
    return INF_ERR_HEADER;
  uint32_t const xlen = (uint32_t)p[10] | ((uint32_t)p[11] << 8);
  if (avail < 12ull + xlen)
    return INF_ERR_INPUT;
  uint32_t at = 12, bsize = 0;
  bool found = false;
  while (at + 4 <= 12 + xlen)
  {
    uint32_t const slen = (uint32_t)p[at + 2] | ((uint32_t)p[at + 3] << 8);
    if (p[at] == 'B' && p[at + 1] == 'C' && slen == 2 && at + 6 <= 12 + xlen)
    {
      bsize = (uint32_t)p[at + 4] | ((uint32_t)p[at + 5] << 8);
      found = true;
    }
    at += 4 + slen;
  }
  if (!found)
    return INF_ERR_HEADER;
  info->header_bytes = 12 + xlen;
  info->block_bytes = bsize + 1;
  if (info->block_bytes < info->header_bytes + 8)
    return INF_ERR_HEADER;
  if (avail < info->block_bytes)
    return INF_ERR_INPUT;
  const uint8_t * t = p + info->block_bytes - 8;
  info->crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
  info->isize = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
  return INF_OK;
}

// Inflates the block at p into out (exactly isize bytes are expected there).  All lanes call it; same result on every lane.
GTB_HD int bgzf_inflate_block(const uint8_t * p, unsigned long long avail, uint8_t * out, InflateTables & T, bool check_crc)
{
  BgzfBlockInfo info{};
  int rc = bgzf_block_info(p, avail, &info);
  if (rc != INF_OK)
    return rc;
  if (info.isize > 65536u)
    return INF_ERR_SIZE;
  uint32_t got = 0;
  rc = inflate_raw(p + info.header_bytes, info.block_bytes - info.header_bytes - 8, out, info.isize, T, &got);
  if (rc != INF_OK)
    return rc;
  if (got != info.isize)
    return INF_ERR_SIZE;
  GTB_INF_SYNC();
  if (check_crc && crc32_coop(out, got) != info.crc)
    return INF_ERR_CRC;
  return INF_OK;
}
} // namespace gtb
