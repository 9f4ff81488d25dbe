// spartan_b200 — zlib stream of `R1CSShape::get_digest` (src/r1cs.rs:154-158: bincode(shape) through flate2's ZlibEncoder at
// Compression::default()).  flate2's default backend is miniz_oxide, the Rust port of miniz; neither is under /root/reference, so this file
// restates miniz's `tdefl` compressor (level 6: 128 probes, lazy parsing, 32 KiB dictionary, 64 KiB LZ code buffer, zlib header 78 9C,
// Adler-32 trailer) decision for decision: hash-chain match finder, lazy-match state machine, block splitting by LZ-buffer fill, raw-block
// fallback, Moffat–Katajainen code lengths on 16-bit counters, length limiting, code-length RLE.  tests/test_deflate.py diffs the output
// bit for bit against the C miniz inside libtorch_cpu.so (mz_compress2 level 6) — the one miniz available in this image; parity with
// miniz_oxide's own bytes is unpinned (no Rust toolchain here).  NIZK::prove absorbs this digest (src/lib.rs:514).
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace sp {

namespace {

enum { LZ_DICT_SIZE = 32768, LZ_DICT_MASK = 32767, MIN_MATCH = 3, MAX_MATCH = 258, LZ_CODE_BUF_SIZE = 64 * 1024, HASH_BITS = 15, HASH_SHIFT = 5, HASH_SIZE = 1 << 15,
       MAX_SYMS0 = 288, MAX_SYMS1 = 32, MAX_SYMS2 = 19, MAX_CODESIZE = 32 };

struct Tdefl {
  // level 6: flags & 0xFFF = 128 probes, lazy parsing
  unsigned max_probes[2] = {1 + (128 + 2) / 3, 1 + ((128 >> 2) + 2) / 3};
  std::vector<uint8_t> dict = std::vector<uint8_t>(LZ_DICT_SIZE + MAX_MATCH - 1, 0);
  std::vector<uint16_t> next = std::vector<uint16_t>(LZ_DICT_SIZE, 0), hash = std::vector<uint16_t>(HASH_SIZE, 0);
  std::vector<uint8_t> lz = std::vector<uint8_t>(LZ_CODE_BUF_SIZE, 0);
  size_t lz_pos = 1, lz_flags = 0;   // m_pLZ_code_buf, m_pLZ_flags as offsets
  unsigned num_flags_left = 8, total_lz_bytes = 0, lz_code_buf_dict_pos = 0, block_index = 0;
  unsigned lookahead_pos = 0, lookahead_size = 0, dict_size = 0, saved_match_dist = 0, saved_match_len = 0, saved_lit = 0;
  uint16_t count[3][MAX_SYMS0] = {}, codes[3][MAX_SYMS0] = {};
  uint8_t sizes[3][MAX_SYMS0] = {};
  uint32_t bit_buffer = 0; unsigned bits_in = 0;
  std::vector<uint8_t> out;
  uint32_t adler = 1;

  void put_bits(uint32_t b, unsigned l) {
    bit_buffer |= b << bits_in; bits_in += l;
    while (bits_in >= 8) { out.push_back((uint8_t)bit_buffer); bit_buffer >>= 8; bits_in -= 8; }
  }
  static unsigned len_sym(unsigned l3) {   // l3 = match_len - 3 -> literal/length symbol
    static const uint16_t base[29] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 255};
    unsigned s = 28;
    while (base[s] > l3) s--;
    return 257 + s;
  }
  static unsigned len_extra(unsigned l3) {
    static const uint8_t ex[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    return ex[len_sym(l3) - 257];
  }
  static void dist_sym(unsigned d1, unsigned& sym, unsigned& extra) {   // d1 = match_dist - 1
    static const uint16_t base[30] = {0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, 12288, 16384, 24576};
    unsigned s = 29;
    while (base[s] > d1) s--;
    sym = s; extra = s < 4 ? 0 : (s - 2) / 2;
  }
  void record_literal(uint8_t lit) {
    total_lz_bytes++;
    lz[lz_pos++] = lit;
    lz[lz_flags] = (uint8_t)(lz[lz_flags] >> 1);
    if (--num_flags_left == 0) { num_flags_left = 8; lz_flags = lz_pos++; }
    count[0][lit]++;
  }
  void record_match(unsigned match_len, unsigned match_dist) {
    total_lz_bytes += match_len;
    lz[lz_pos] = (uint8_t)(match_len - MIN_MATCH);
    match_dist -= 1;
    lz[lz_pos + 1] = (uint8_t)(match_dist & 0xFF);
    lz[lz_pos + 2] = (uint8_t)(match_dist >> 8);
    lz_pos += 3;
    lz[lz_flags] = (uint8_t)((lz[lz_flags] >> 1) | 0x80);
    if (--num_flags_left == 0) { num_flags_left = 8; lz_flags = lz_pos++; }
    unsigned s, e;
    dist_sym(match_dist, s, e);
    count[1][s]++;
    count[0][len_sym(match_len - MIN_MATCH)]++;
  }
  void find_match(unsigned la_pos, unsigned max_dist, unsigned max_match_len, unsigned& match_dist, unsigned& match_len_io) {
    unsigned dist = 0, pos = la_pos & LZ_DICT_MASK, match_len = match_len_io, probe_pos = pos, next_probe_pos, probe_len;
    unsigned num_probes_left = max_probes[match_len >= 32];
    const uint8_t* s = dict.data() + pos;
    uint8_t c0 = dict[pos + match_len], c1 = dict[pos + match_len - 1];
    if (max_match_len <= match_len) return;
    for (;;) {
      for (;;) {
        if (--num_probes_left == 0) return;
        bool hit = false;
        for (int rep = 0; rep < 3; rep++) {
          next_probe_pos = next[probe_pos];
          if (!next_probe_pos || (dist = (uint16_t)(la_pos - next_probe_pos)) > max_dist) return;
          probe_pos = next_probe_pos & LZ_DICT_MASK;
          if (dict[probe_pos + match_len] == c0 && dict[probe_pos + match_len - 1] == c1) { hit = true; break; }
        }
        if (hit) break;
      }
      if (!dist) break;
      const uint8_t *p = s, *q = dict.data() + probe_pos;
      for (probe_len = 0; probe_len < max_match_len; probe_len++) if (*p++ != *q++) break;
      if (probe_len > match_len) {
        match_dist = dist;
        if ((match_len_io = match_len = probe_len) == max_match_len) return;
        c0 = dict[pos + match_len]; c1 = dict[pos + match_len - 1];
      }
    }
  }
  // ---- Huffman tables
  struct SymFreq { uint16_t key, sym; };
  static void minimum_redundancy(SymFreq* A, int n) {   // Moffat & Katajainen, in place, on 16-bit keys (wrap-around included, as in miniz)
    int root, leaf, nxt, avbl, used, dpth;
    if (n == 0) return;
    if (n == 1) { A[0].key = 1; return; }
    A[0].key = (uint16_t)(A[0].key + A[1].key); root = 0; leaf = 2;
    for (nxt = 1; nxt < n - 1; nxt++) {
      if (leaf >= n || A[root].key < A[leaf].key) { A[nxt].key = A[root].key; A[root++].key = (uint16_t)nxt; } else A[nxt].key = A[leaf++].key;
      if (leaf >= n || (root < nxt && A[root].key < A[leaf].key)) { A[nxt].key = (uint16_t)(A[nxt].key + A[root].key); A[root++].key = (uint16_t)nxt; }
      else A[nxt].key = (uint16_t)(A[nxt].key + A[leaf++].key);
    }
    A[n - 2].key = 0;
    for (nxt = n - 3; nxt >= 0; nxt--) A[nxt].key = (uint16_t)(A[A[nxt].key].key + 1);
    avbl = 1; used = dpth = 0; root = n - 2; nxt = n - 1;
    while (avbl > 0) {
      while (root >= 0 && (int)A[root].key == dpth) { used++; root--; }
      while (avbl > used) { A[nxt--].key = (uint16_t)dpth; avbl--; }
      avbl = 2 * used; dpth++; used = 0;
    }
  }
  static void enforce_max_code_size(int* num_codes, int code_list_len, int max_code_size) {
    if (code_list_len <= 1) return;
    uint32_t total = 0;
    for (int i = max_code_size + 1; i <= MAX_CODESIZE; i++) num_codes[max_code_size] += num_codes[i];
    for (int i = max_code_size; i > 0; i--) total += ((uint32_t)num_codes[i]) << (max_code_size - i);
    while (total != (1u << max_code_size)) {
      num_codes[max_code_size]--;
      for (int i = max_code_size - 1; i > 0; i--) if (num_codes[i]) { num_codes[i]--; num_codes[i + 1] += 2; break; }
      total--;
    }
  }
  void optimize_table(int t, int table_len, int code_size_limit, bool static_table) {
    int num_codes[1 + MAX_CODESIZE] = {};
    unsigned next_code[MAX_CODESIZE + 1];
    if (static_table) { for (int i = 0; i < table_len; i++) num_codes[sizes[t][i]]++; }
    else {
      SymFreq syms[MAX_SYMS0];
      int used = 0;
      for (int i = 0; i < table_len; i++) if (count[t][i]) { syms[used].key = count[t][i]; syms[used++].sym = (uint16_t)i; }
      std::stable_sort(syms, syms + used, [](const SymFreq& a, const SymFreq& b) { return a.key < b.key; });   // = miniz's two-pass LSD radix sort
      minimum_redundancy(syms, used);
      for (int i = 0; i < used; i++) num_codes[syms[i].key]++;
      enforce_max_code_size(num_codes, used, code_size_limit);
      memset(sizes[t], 0, sizeof sizes[t]); memset(codes[t], 0, sizeof codes[t]);
      for (int i = 1, j = used; i <= code_size_limit; i++) for (int l = num_codes[i]; l > 0; l--) sizes[t][syms[--j].sym] = (uint8_t)i;
    }
    next_code[1] = 0;
    for (int j = 0, i = 2; i <= code_size_limit; i++) next_code[i] = j = ((j + num_codes[i - 1]) << 1);
    for (int i = 0; i < table_len; i++) {
      unsigned rev = 0, code, cs = sizes[t][i];
      if (!cs) continue;
      code = next_code[cs]++;
      for (unsigned l = cs; l > 0; l--, code >>= 1) rev = (rev << 1) | (code & 1);
      codes[t][i] = (uint16_t)rev;
    }
  }
  void start_static_block() {
    uint8_t* p = sizes[0];
    int i = 0;
    for (; i <= 143; ++i) p[i] = 8;
    for (; i <= 255; ++i) p[i] = 9;
    for (; i <= 279; ++i) p[i] = 7;
    for (; i <= 287; ++i) p[i] = 8;
    memset(sizes[1], 5, 32);
    optimize_table(0, 288, 15, true);
    optimize_table(1, 32, 15, true);
    put_bits(1, 2);
  }
  void start_dynamic_block() {
    static const uint8_t swizzle[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t to_pack[MAX_SYMS0 + MAX_SYMS1], packed[MAX_SYMS0 + MAX_SYMS1], prev = 0xFF;
    count[0][256] = 1;
    optimize_table(0, MAX_SYMS0, 15, false);
    optimize_table(1, MAX_SYMS1, 15, false);
    int num_lit, num_dist;
    for (num_lit = 286; num_lit > 257; num_lit--) if (sizes[0][num_lit - 1]) break;
    for (num_dist = 30; num_dist > 1; num_dist--) if (sizes[1][num_dist - 1]) break;
    memcpy(to_pack, sizes[0], num_lit);
    memcpy(to_pack + num_lit, sizes[1], num_dist);
    unsigned total = num_lit + num_dist, np = 0, rle_z = 0, rle_rep = 0;
    memset(count[2], 0, sizeof(count[2][0]) * MAX_SYMS2);
    auto rle_prev = [&]() {
      if (rle_rep) {
        if (rle_rep < 3) { count[2][prev] = (uint16_t)(count[2][prev] + rle_rep); while (rle_rep--) packed[np++] = prev; }
        else { count[2][16] = (uint16_t)(count[2][16] + 1); packed[np++] = 16; packed[np++] = (uint8_t)(rle_rep - 3); }
        rle_rep = 0;
      }
    };
    auto rle_zero = [&]() {
      if (rle_z) {
        if (rle_z < 3) { count[2][0] = (uint16_t)(count[2][0] + rle_z); while (rle_z--) packed[np++] = 0; }
        else if (rle_z <= 10) { count[2][17] = (uint16_t)(count[2][17] + 1); packed[np++] = 17; packed[np++] = (uint8_t)(rle_z - 3); }
        else { count[2][18] = (uint16_t)(count[2][18] + 1); packed[np++] = 18; packed[np++] = (uint8_t)(rle_z - 11); }
        rle_z = 0;
      }
    };
    for (unsigned i = 0; i < total; i++) {
      uint8_t cs = to_pack[i];
      if (!cs) {
        rle_prev();
        if (++rle_z == 138) rle_zero();
      } else {
        rle_zero();
        if (cs != prev) { rle_prev(); count[2][cs] = (uint16_t)(count[2][cs] + 1); packed[np++] = cs; }
        else if (++rle_rep == 6) rle_prev();
      }
      prev = cs;
    }
    if (rle_rep) rle_prev(); else rle_zero();
    optimize_table(2, MAX_SYMS2, 7, false);
    put_bits(2, 2);
    put_bits(num_lit - 257, 5);
    put_bits(num_dist - 1, 5);
    int nbl;
    for (nbl = 18; nbl >= 0; nbl--) if (sizes[2][swizzle[nbl]]) break;
    nbl = std::max(4, nbl + 1);
    put_bits(nbl - 4, 4);
    for (int i = 0; i < nbl; i++) put_bits(sizes[2][swizzle[i]], 3);
    for (unsigned k = 0; k < np;) {
      unsigned code = packed[k++];
      put_bits(codes[2][code], sizes[2][code]);
      if (code >= 16) put_bits(packed[k++], "\02\03\07"[code - 16]);
    }
  }
  void compress_lz_codes() {
    unsigned flags = 1;
    for (size_t p = 0; p < lz_pos; flags >>= 1) {
      if (flags == 1) flags = lz[p++] | 0x100;
      if (flags & 1) {
        unsigned ml = lz[p], md = lz[p + 1] | (lz[p + 2] << 8);
        p += 3;
        unsigned ls = len_sym(ml), le = len_extra(ml), ds, de;
        put_bits(codes[0][ls], sizes[0][ls]);
        put_bits(ml & ((1u << le) - 1), le);
        dist_sym(md, ds, de);
        put_bits(codes[1][ds], sizes[1][ds]);
        put_bits(md & ((1u << de) - 1), de);
      } else {
        unsigned lit = lz[p++];
        put_bits(codes[0][lit], sizes[0][lit]);
      }
    }
    put_bits(codes[0][256], sizes[0][256]);
  }
  void compress_block(bool static_block) {
    if (static_block) start_static_block(); else start_dynamic_block();
    compress_lz_codes();
  }
  void flush_block(bool finish) {
    lz[lz_flags] = (uint8_t)(lz[lz_flags] >> num_flags_left);
    lz_pos -= (num_flags_left == 8);
    if (!block_index) { put_bits(0x78, 8); put_bits(0x9C, 8); }   // zlib header: 32 KiB window, FLEVEL 2 (the level whose probe count is 128)
    put_bits(finish ? 1 : 0, 1);
    const size_t saved_out = out.size();
    const uint32_t saved_bit_buf = bit_buffer; const unsigned saved_bits_in = bits_in;
    compress_block(total_lz_bytes < 48);
    // if the block got expanded, send it raw instead (only possible while its bytes are still in the dictionary)
    if (total_lz_bytes && (out.size() - saved_out + 1U) >= total_lz_bytes && (lookahead_pos - lz_code_buf_dict_pos) <= dict_size) {
      out.resize(saved_out); bit_buffer = saved_bit_buf; bits_in = saved_bits_in;
      put_bits(0, 2);
      if (bits_in) put_bits(0, 8 - bits_in);
      put_bits(total_lz_bytes & 0xFFFF, 16);
      put_bits((total_lz_bytes ^ 0xFFFF) & 0xFFFF, 16);
      for (unsigned i = 0; i < total_lz_bytes; ++i) put_bits(dict[(lz_code_buf_dict_pos + i) & LZ_DICT_MASK], 8);
    }
    if (finish) {
      if (bits_in) put_bits(0, 8 - bits_in);
      uint32_t a = adler;
      for (int i = 0; i < 4; i++) { put_bits((a >> 24) & 0xFF, 8); a <<= 8; }
    }
    memset(count[0], 0, sizeof count[0]); memset(count[1], 0, sizeof count[1]);
    lz_pos = 1; lz_flags = 0; num_flags_left = 8;
    lz_code_buf_dict_pos += total_lz_bytes; total_lz_bytes = 0; block_index++;
  }
  void compress(const uint8_t* src, size_t src_left) {   // the whole input with TDEFL_FINISH
    {  // Adler-32
      uint32_t s1 = 1, s2 = 0;
      const uint8_t* p = src; size_t n = src_left;
      while (n) { size_t blk = n < 5552 ? n : 5552; for (size_t i = 0; i < blk; i++) { s1 += p[i]; s2 += s1; } s1 %= 65521u; s2 %= 65521u; p += blk; n -= blk; }
      adler = (s2 << 16) | s1;
    }
    while (src_left || lookahead_size) {
      if ((lookahead_size + dict_size) >= (MIN_MATCH - 1)) {
        unsigned dst_pos = (lookahead_pos + lookahead_size) & LZ_DICT_MASK, ins_pos = lookahead_pos + lookahead_size - 2;
        unsigned h = (dict[ins_pos & LZ_DICT_MASK] << HASH_SHIFT) ^ dict[(ins_pos + 1) & LZ_DICT_MASK];
        unsigned n = (unsigned)std::min<size_t>(src_left, MAX_MATCH - lookahead_size);
        src_left -= n; lookahead_size += n;
        for (unsigned k = 0; k < n; k++) {
          uint8_t c = *src++;
          dict[dst_pos] = c;
          if (dst_pos < (MAX_MATCH - 1)) dict[LZ_DICT_SIZE + dst_pos] = c;
          h = ((h << HASH_SHIFT) ^ c) & (HASH_SIZE - 1);
          next[ins_pos & LZ_DICT_MASK] = hash[h];
          hash[h] = (uint16_t)ins_pos;
          dst_pos = (dst_pos + 1) & LZ_DICT_MASK;
          ins_pos++;
        }
      } else {
        while (src_left && lookahead_size < MAX_MATCH) {
          uint8_t c = *src++;
          unsigned dst_pos = (lookahead_pos + lookahead_size) & LZ_DICT_MASK;
          src_left--;
          dict[dst_pos] = c;
          if (dst_pos < (MAX_MATCH - 1)) dict[LZ_DICT_SIZE + dst_pos] = c;
          if ((++lookahead_size + dict_size) >= MIN_MATCH) {
            unsigned ins_pos = lookahead_pos + (lookahead_size - 1) - 2;
            unsigned h = ((dict[ins_pos & LZ_DICT_MASK] << (HASH_SHIFT * 2)) ^ (dict[(ins_pos + 1) & LZ_DICT_MASK] << HASH_SHIFT) ^ c) & (HASH_SIZE - 1);
            next[ins_pos & LZ_DICT_MASK] = hash[h];
            hash[h] = (uint16_t)ins_pos;
          }
        }
      }
      dict_size = std::min<unsigned>(LZ_DICT_SIZE - lookahead_size, dict_size);
      // (flush == FINISH: never wait for more input)
      unsigned len_to_move = 1, cur_match_dist = 0, cur_match_len = saved_match_len ? saved_match_len : (MIN_MATCH - 1);
      const unsigned cur_pos = lookahead_pos & LZ_DICT_MASK;
      find_match(lookahead_pos, dict_size, lookahead_size, cur_match_dist, cur_match_len);
      if ((cur_match_len == MIN_MATCH && cur_match_dist >= 8U * 1024U) || cur_pos == cur_match_dist) cur_match_dist = cur_match_len = 0;
      if (saved_match_len) {
        if (cur_match_len > saved_match_len) {
          record_literal((uint8_t)saved_lit);
          if (cur_match_len >= 128) { record_match(cur_match_len, cur_match_dist); saved_match_len = 0; len_to_move = cur_match_len; }
          else { saved_lit = dict[cur_pos]; saved_match_dist = cur_match_dist; saved_match_len = cur_match_len; }
        } else {
          record_match(saved_match_len, saved_match_dist);
          len_to_move = saved_match_len - 1; saved_match_len = 0;
        }
      } else if (!cur_match_dist) record_literal(dict[cur_pos]);
      else if (cur_match_len >= 128) { record_match(cur_match_len, cur_match_dist); len_to_move = cur_match_len; }
      else { saved_lit = dict[cur_pos]; saved_match_dist = cur_match_dist; saved_match_len = cur_match_len; }
      lookahead_pos += len_to_move;
      lookahead_size -= len_to_move;
      dict_size = std::min<unsigned>(dict_size + len_to_move, LZ_DICT_SIZE);
      if (lz_pos > (size_t)(LZ_CODE_BUF_SIZE - 8) || (total_lz_bytes > 31 * 1024 && ((((unsigned)lz_pos * 115) >> 7) >= total_lz_bytes))) flush_block(false);
    }
    flush_block(true);
  }
};

}  // namespace

// zlib stream, bit-identical to miniz's mz_compress2(.., level 6) on the same bytes
std::vector<uint8_t> miniz_zlib_level6(const uint8_t* data, size_t len) {
  Tdefl t;
  t.out.reserve(len / 2 + 64);
  t.compress(data, len);
  return std::move(t.out);
}

}  // namespace sp
