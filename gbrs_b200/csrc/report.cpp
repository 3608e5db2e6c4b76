// Native writer for the report tables (reference: the per-row python loops of EMfactory.report_read_counts /
// report_depths, src/gbrs/emase/EMfactory.py:318-331 and :366-380, and AlignmentPropertyMatrix.report_alignment_counts,
// src/gbrs/emase/AlignmentPropertyMatrix.py:442-459).  Every value is printed exactly like python's
// `str(numpy.float64)` / `repr(float)`: shortest round-trip digits, fixed notation for 1e-4 <= |x| < 1e16, otherwise
// scientific with a two-digit exponent, always a ".0" on integers.  With the EM itself down to milliseconds, formatting
// ~10^6 numbers per table in python had become the slowest part of `quantify` after packing.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "gbrs_em.h"

void gbrs_set_error(const std::string& s);

namespace {

// append repr(float(x)) to out
void append_py_float(std::string& out, double x) {
  if (std::isnan(x)) { out += "nan"; return; }
  if (std::isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);  // shortest round-trip digits
  const char* p = buf;
  const char* end = res.ptr;
  if (*p == '-') { out += '-'; ++p; }
  char digits[32];
  int nd = 0;
  while (p < end && *p != 'e') {
    if (*p != '.') digits[nd++] = *p;
    ++p;
  }
  int exp10 = 0;
  if (p < end) {  // "e[+-]dd"
    ++p;
    bool neg = false;
    if (*p == '+') ++p; else if (*p == '-') { neg = true; ++p; }
    while (p < end) exp10 = exp10 * 10 + (*p++ - '0');
    if (neg) exp10 = -exp10;
  }
  while (nd > 1 && digits[nd - 1] == '0') --nd;  // to_chars never pads, but be safe
  const int decpt = exp10 + 1;                   // value = 0.d1 d2 ... x 10^decpt
  if (decpt > -4 && decpt <= 16) {
    if (decpt <= 0) {
      out += "0.";
      out.append((size_t) -decpt, '0');
      out.append(digits, (size_t) nd);
    } else if (decpt >= nd) {
      out.append(digits, (size_t) nd);
      out.append((size_t) (decpt - nd), '0');
      out += ".0";
    } else {
      out.append(digits, (size_t) decpt);
      out += '.';
      out.append(digits + decpt, (size_t) (nd - decpt));
    }
  } else {
    out += digits[0];
    if (nd > 1) {
      out += '.';
      out.append(digits + 1, (size_t) (nd - 1));
    }
    out += 'e';
    int e = decpt - 1;
    out += e < 0 ? '-' : '+';
    if (e < 0) e = -e;
    char eb[8];
    int ne = 0;
    do { eb[ne++] = (char) ('0' + e % 10); e /= 10; } while (e);
    if (ne < 2) eb[ne++] = '0';
    while (ne) out += eb[--ne];
  }
}

}  // namespace

extern "C" int gbrs_format_double(double x, char* out, int32_t cap) {
  std::string s;
  append_py_float(s, x);
  if ((int) s.size() + 1 > cap) return GBRS_E_ARG;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int) s.size();
}

extern "C" int gbrs_write_table(const char* path, const char* header, const char* const* names, int64_t n_rows,
                                const double* data, int32_t n_cols, const char* const* notes, const int64_t* order,
                                int32_t append) {
  if (!path || !names || !data || n_rows < 0 || n_cols < 0) { gbrs_set_error("gbrs_write_table: bad argument"); return GBRS_E_ARG; }
  std::FILE* fh = std::fopen(path, append ? "ab" : "wb");
  if (!fh) { gbrs_set_error(std::string("gbrs_write_table: cannot open ") + path); return GBRS_E_ARG; }
  if (header) std::fputs(header, fh);
  const int64_t chunk = 2048;
  const int64_t n_chunks = (n_rows + chunk - 1) / chunk;
  std::vector<std::string> parts((size_t) n_chunks);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < n_chunks; ++c) {
    std::string& s = parts[(size_t) c];
    s.reserve((size_t) chunk * (24 + 20 * (size_t) n_cols));
    const int64_t lo = c * chunk, hi = lo + chunk < n_rows ? lo + chunk : n_rows;
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t r = order ? order[i] : i;
      s += names[r];
      for (int32_t k = 0; k < n_cols; ++k) {
        s += '\t';
        append_py_float(s, data[(size_t) k * (size_t) n_rows + (size_t) r]);  // data is [n_cols][n_rows]
      }
      if (notes) {
        s += '\t';
        s += notes[r];
      }
      s += '\n';
    }
  }
  bool ok = true;
  for (const std::string& s : parts) ok = ok && std::fwrite(s.data(), 1, s.size(), fh) == s.size();
  ok = (std::fclose(fh) == 0) && ok;
  if (!ok) { gbrs_set_error(std::string("gbrs_write_table: write failed for ") + path); return GBRS_E_ARG; }
  return GBRS_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Length table of EMfactory.prepare (src/gbrs/emase/EMfactory.py:60-89): lines `locus_hap<TAB>length` (or
// `locus<TAB>length` with one haplotype) -> out[locus][hap] = max(length - read_length + 1, 1).  The reference walks
// the file in a python loop (about a second per 10^6 lines).  Only well-formed lines are handled here; at the first
// line that is not (unknown name, key that does not split into exactly two parts at '_', anything python's float()
// might still accept but strtod in the C locale does not take whole) the function stops and reports the line, and the
// caller re-parses with the reference-style loop so that errors surface exactly as the reference raises them.
// ---------------------------------------------------------------------------------------------------------------------
#include <fstream>
#include <iterator>
#include <unordered_map>

extern "C" int gbrs_parse_lengths(const char* path, const char* const* lnames, int64_t n_loci, const char* const* hnames,
                                  int32_t n_haps, double read_length, double* out, int64_t* bad_line) {
  if (!path || !lnames || !out || n_loci < 1 || n_haps < 1 || (n_haps > 1 && !hnames) || !bad_line) {
    gbrs_set_error("gbrs_parse_lengths: bad argument"); return GBRS_E_ARG;
  }
  *bad_line = 0;
  std::ifstream fh(path, std::ios::binary);
  if (!fh) { gbrs_set_error(std::string("gbrs_parse_lengths: cannot open ") + path); return GBRS_E_ARG; }
  std::string data((std::istreambuf_iterator<char>(fh)), std::istreambuf_iterator<char>());
  std::unordered_map<std::string_view, int64_t> lid, hid;
  lid.reserve((size_t) n_loci * 2);
  for (int64_t i = 0; i < n_loci; ++i) lid[std::string_view(lnames[i])] = i;  // later duplicates win, like dict(zip(...))
  for (int32_t h = 0; h < n_haps && hnames; ++h) hid[std::string_view(hnames[h])] = h;
  for (int64_t i = 0; i < n_loci * n_haps; ++i) out[i] = 0.0;
  const char* p = data.data();
  const char* end = p + data.size();
  int64_t line_no = 0;
  auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
  while (p < end) {
    ++line_no;
    const char* eol = static_cast<const char*>(std::memchr(p, '\n', (size_t) (end - p)));
    const char* next = eol ? eol + 1 : end;
    const char* le = eol ? eol : end;
    while (le > p && is_space(le[-1])) --le;  // curline.rstrip()
    const char* tab = static_cast<const char*>(std::memchr(p, '\t', (size_t) (le - p)));
    if (!tab) { *bad_line = line_no; return GBRS_OK; }
    const char* v0 = tab + 1;
    const char* v1 = static_cast<const char*>(std::memchr(v0, '\t', (size_t) (le - v0)));
    if (!v1) v1 = le;
    // the number: plain decimal / exponent notation only, consumed whole (surrounding blanks allowed, as float() does)
    while (v0 < v1 && is_space(*v0)) ++v0;
    while (v1 > v0 && is_space(v1[-1])) --v1;
    double len = 0.0;
    auto res = std::from_chars(v0, v1, len);
    if (v0 == v1 || res.ec != std::errc() || res.ptr != v1) { *bad_line = line_no; return GBRS_OK; }
    std::string_view key(p, (size_t) (tab - p));
    int64_t li = -1, hi = 0;
    if (n_haps > 1) {
      const size_t us = key.find('_');
      if (us == std::string_view::npos || key.find('_', us + 1) != std::string_view::npos) { *bad_line = line_no; return GBRS_OK; }
      auto a = lid.find(key.substr(0, us));
      auto b = hid.find(key.substr(us + 1));
      if (a == lid.end() || b == hid.end()) { *bad_line = line_no; return GBRS_OK; }
      li = a->second;
      hi = b->second;
    } else {
      auto a = lid.find(key);
      if (a == lid.end()) { *bad_line = line_no; return GBRS_OK; }
      li = a->second;
    }
    const double eff = len - read_length + 1.0;
    out[li * n_haps + hi] = eff > 1.0 ? eff : 1.0;  // max(float(item[1]) - read_length + 1.0, 1.0)
    p = next;
  }
  return GBRS_OK;
}
