// Host-side builder of the TILE layout read by the fused model-4 kernel (k_tile_em, em_kernels.cu); the format is
// documented in include/gbrs_em.h.  Input: a packed shard (pack.cpp) -- class-major pair words and class counts.
//
// What it replaces in the reference: the same storage walk as pack.cpp (Sparse3DMatrix.py:26-66, the per-iteration
// passes of AlignmentPropertyMatrix.normalize_reads / sum, :275-370), re-laid so that one thread block can run the
// E-step and the M-step of a contiguous range of classes entirely out of shared memory.
//
//   1. classes are ordered by their smallest locus (all widths together) and cut greedily into tiles bounded by class,
//      locus, pair, entry and item caps -- a tile's classes touch few distinct loci, because classes that share their
//      smallest locus share most of the others (reads multi-map within a gene family);
//   2. every tile gets its sorted locus list (pair words then carry an 8-bit LOCAL locus), its classes sorted by width
//      (descending), the pair words as planes, and its own locus-major copy: local class ids grouped by
//      (local locus, nibble bucket), cut into work items;
//   3. the (tile, locus) output slots are numbered locus-major, so that the locus kernel finds the partial sums of a
//      locus in consecutive slots, in tile order (fixed summation order: results are bit-reproducible).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "pack_internal.h"

struct gbrs_tiles {
  gbrs_tiles_info info{};
  bigvec<uint8_t> blob;
  std::vector<uint32_t> tile_desc, locus_desc;
};

namespace {

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct TileStat {
  int64_t first = 0;  // index into `order` of the tile's first class
  int32_t n_classes = 0, n_loci = 0, n_pairs = 0, n_entries = 0, n_items = 0, n_planes = 0, n_runs = 0, n_slices = 0;
  int64_t a_bytes = 0, b_bytes = 0, blob_off = 0;
  uint32_t off_loci = 0, off_slots = 0, off_a[4] = {0, 0, 0, 0}, off_b[4] = {0, 0, 0, 0};
};

// entries a pair word contributes to the tile's locus-major copy
inline int entries_of(uint32_t mask, uint32_t full) { return mask == full ? 1 : ((mask & 15u) != 0) + ((mask >> 4) != 0); }

}  // namespace

extern "C" int gbrs_tiles_create(gbrs_pack_t P, const gbrs_tiles_params* prm, gbrs_tiles_t* out) {
  if (!P || !out) { gbrs_set_error("gbrs_tiles_create: null argument"); return GBRS_E_ARG; }
  gbrs_tiles_params q{};
  if (prm) q = *prm;
  const int maxC = q.max_classes > 0 ? q.max_classes : 128;
  const int maxL = q.max_loci > 0 ? q.max_loci : 16;
  const int maxP = q.max_pairs > 0 ? q.max_pairs : 65535;   // pair words and entries are streamed, not staged:
  const int maxE = q.max_entries > 0 ? q.max_entries : 65535; // no cap is needed beyond the 16-bit fields
  const int maxI = q.max_items > 0 ? q.max_items : 192;
  const int ilen = q.item_len > 0 ? q.item_len : 16;
  if (maxC > 2047 || maxL > 128 || maxE > 65535 || maxP > 65535 || maxI > 65535 || ilen > 16 || maxP < maxL || maxE < 2 * maxL || maxI < 2 * maxL) {
    gbrs_set_error("gbrs_tiles_create: tile caps out of range"); return GBRS_E_ARG;
  }
  const OmpThreadsGuard omp_guard(pack_threads());
  const int T = P->T;
  const int64_t n = P->info.n_classes;
  const uint32_t* rowptr = P->rowptr.data();
  const uint32_t* pairs = P->pairs.data();
  const double* count = P->count.data();
  const uint32_t full = (1u << P->H) - 1u;  // a pair with this mask hits every haplotype: one entry in bucket 0
  try {
    // ---- 1. order by smallest locus ------------------------------------------------------------------------------
    bigvec<uint32_t> minloc((size_t) n);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n; ++c) {
      uint32_t m = 0xFFFFFFFFu;
      for (uint32_t p = rowptr[c]; p < rowptr[c + 1]; ++p) m = std::min(m, pairs[p] & 0xFFFFFFu);
      minloc[c] = m;
    }
    std::vector<int64_t> start((size_t) T + 1, 0);
    for (int64_t c = 0; c < n; ++c) ++start[minloc[c] + 1];
    for (int t = 0; t < T; ++t) start[t + 1] += start[t];
    bigvec<uint32_t> order((size_t) n);
    {
      std::vector<int64_t> cur(start.begin(), start.end() - 1);
      for (int64_t c = 0; c < n; ++c) order[cur[minloc[c]]++] = (uint32_t) c;
    }
    bigvec<uint32_t>().swap(minloc);

    // ---- 2. greedy cut ---------------------------------------------------------------------------------------------
    std::vector<TileStat> tiles;
    {
      std::vector<int32_t> locus_stamp((size_t) T, -1);
      bigvec<int32_t> key_stamp;
      bigvec<uint16_t> key_count;
      par_fill(key_stamp, (size_t) T * 32, -1);
      par_fill(key_count, (size_t) T * 32, (uint16_t) 0);
      TileStat cur;
      int32_t id = 0;
      auto close = [&]() { tiles.push_back(cur); ++id; };
      for (int64_t i = 0; i < n; ++i) {
        const uint32_t c = order[i];
        const uint32_t b = rowptr[c], e = rowptr[c + 1];
        const int k = (int) (e - b);
        if (k > maxL) {
          gbrs_set_error("gbrs_tiles_create: a class touches more loci than a tile may hold");
          return GBRS_E_LIMIT;
        }
        for (int attempt = 0; attempt < 2; ++attempt) {
          // what adding the class would cost in the current tile
          int new_loci = 0, ent = 0, items = 0, runs = 0;
          for (uint32_t p = b; p < e; ++p) {
            const uint32_t w = pairs[p], t = w & 0xFFFFFFu, m = w >> 24;
            new_loci += locus_stamp[t] != id;
            uint32_t keys[2];
            int nk = 0;
            if (m == full) keys[nk++] = t * 32;
            else {
              if (m & 15u) keys[nk++] = t * 32 + (m & 15u);
              if (m >> 4) keys[nk++] = t * 32 + 16 + (m >> 4);
            }
            ent += nk;
            for (int j = 0; j < nk; ++j) {
              const int cnt = key_stamp[keys[j]] == id ? key_count[keys[j]] : 0;
              items += cnt % ilen == 0;  // (two keys of one class never coincide: different buckets or loci)
              runs += cnt == 0;
            }
          }
          const bool fits = cur.n_classes + 1 <= maxC && cur.n_loci + new_loci <= maxL && cur.n_pairs + k <= maxP &&
                            cur.n_entries + ent <= maxE && cur.n_items + items <= maxI;
          if (!fits && cur.n_classes > 0) {
            close();
            cur = TileStat();
            cur.first = i;
            continue;  // re-evaluate against the empty tile (stamps of the old tile no longer match `id`)
          }
          for (uint32_t p = b; p < e; ++p) {
            const uint32_t w = pairs[p], t = w & 0xFFFFFFu, m = w >> 24;
            locus_stamp[t] = id;
            uint32_t keys[2];
            int nk = 0;
            if (m == full) keys[nk++] = t * 32;
            else {
              if (m & 15u) keys[nk++] = t * 32 + (m & 15u);
              if (m >> 4) keys[nk++] = t * 32 + 16 + (m >> 4);
            }
            for (int j = 0; j < nk; ++j) {
              if (key_stamp[keys[j]] != id) { key_stamp[keys[j]] = id; key_count[keys[j]] = 0; }
              ++key_count[keys[j]];
            }
          }
          cur.n_classes += 1;
          cur.n_loci += new_loci;
          cur.n_pairs += k;
          cur.n_entries += ent;
          cur.n_items += items;
          cur.n_runs += runs;
          cur.n_planes = std::max(cur.n_planes, k);
          break;
        }
      }
      if (cur.n_classes > 0) close();
    }
    const int64_t n_tiles = (int64_t) tiles.size();

    // ---- 3. build every tile into its own byte vector (parallel) -----------------------------------------------------
    auto* R = new gbrs_tiles();
    gbrs_tiles_info& info = R->info;
    std::vector<std::vector<uint8_t>> bytes((size_t) n_tiles);
    static const int kSliceLen[17] = {0, 1, 2, 3, 4, 6, 6, 8, 8, 12, 12, 12, 12, 16, 16, 16, 16};  // padded slice lengths
    int failed = 0;
#pragma omp parallel
    {
      std::vector<uint32_t> loci, cls, keyed, kitems, vis;
#pragma omp for schedule(dynamic, 8)
      for (int64_t u = 0; u < n_tiles; ++u) {
        TileStat& t = tiles[u];
        // locus list
        loci.clear();
        for (int64_t i = t.first; i < t.first + t.n_classes; ++i) {
          const uint32_t c = order[i];
          for (uint32_t p = rowptr[c]; p < rowptr[c + 1]; ++p) loci.push_back(pairs[p] & 0xFFFFFFu);
        }
        std::sort(loci.begin(), loci.end());
        loci.erase(std::unique(loci.begin(), loci.end()), loci.end());
        if ((int) loci.size() != t.n_loci) { failed = 1; continue; }
        // classes by descending width (stable: smallest-locus order is kept inside a width)
        cls.assign(order.begin() + t.first, order.begin() + t.first + t.n_classes);
        std::stable_sort(cls.begin(), cls.end(), [&](uint32_t x, uint32_t y) {
          return rowptr[x + 1] - rowptr[x] > rowptr[y + 1] - rowptr[y];
        });
        std::vector<uint32_t> nplane((size_t) t.n_planes, 0), plane_off((size_t) t.n_planes + 1, 0);
        for (int j = 0; j < t.n_classes; ++j) {
          const int k = (int) (rowptr[cls[j] + 1] - rowptr[cls[j]]);
          for (int p = 0; p < k; ++p) ++nplane[p];
        }
        for (int p = 0; p < t.n_planes; ++p) plane_off[p + 1] = plane_off[p] + (nplane[p] + 3u) / 4u * 4u;
        // part A geometry
        const int64_t off_loci = GBRS_TH_WORDS * 4;
        const int64_t off_slots = align_up(off_loci + 4 * (int64_t) t.n_loci, 16);
        const int64_t off_nplane = align_up(off_slots + 4 * (int64_t) t.n_loci, 16);
        const int64_t off_count = align_up(off_nplane + 2 * (int64_t) t.n_planes, 16);
        const int64_t off_pairs = align_up(off_count + 8 * (int64_t) t.n_classes, 16);
        const int64_t a_bytes = align_up(off_pairs + 2 * (int64_t) plane_off[t.n_planes], 16);
        // the tile's locus-major copy: (key, class) sorted, cut into items, items sorted by length
        keyed.clear();
        std::vector<uint16_t> pw((size_t) plane_off[t.n_planes], 0);
        for (int j = 0; j < t.n_classes; ++j) {
          const uint32_t b = rowptr[cls[j]], e = rowptr[cls[j] + 1];
          for (uint32_t p = b; p < e; ++p) {
            const uint32_t w = pairs[p], loc = w & 0xFFFFFFu, m = w >> 24;
            const uint32_t l = (uint32_t) (std::lower_bound(loci.begin(), loci.end(), loc) - loci.begin());
            pw[plane_off[p - b] + j] = (uint16_t) ((l << 8) | m);
            if (m == full) keyed.push_back(((l * 32u) << 16) | (uint32_t) j);
            else {
              if (m & 15u) keyed.push_back(((l * 32u + (m & 15u)) << 16) | (uint32_t) j);
              if (m >> 4) keyed.push_back(((l * 32u + 16u + (m >> 4)) << 16) | (uint32_t) j);
            }
          }
        }
        if ((int) keyed.size() != t.n_entries) { failed = 1; continue; }
        std::sort(keyed.begin(), keyed.end());
        kitems.clear();  // key order: first entry | (len - 1) << 16
        std::vector<uint16_t> run_key, run_first;
        for (size_t i = 0; i < keyed.size();) {
          const uint32_t key = keyed[i] >> 16;
          size_t j = i;
          while (j < keyed.size() && (keyed[j] >> 16) == key) ++j;
          run_key.push_back((uint16_t) key);
          run_first.push_back((uint16_t) kitems.size());
          for (size_t st = i; st < j; st += (size_t) ilen) {
            const uint32_t len = (uint32_t) std::min<size_t>((size_t) ilen, j - st);
            kitems.push_back((uint32_t) st | ((len - 1) << 16));
          }
          i = j;
        }
        const int ni = (int) kitems.size(), nr = (int) run_key.size();
        if (ni != t.n_items || nr != t.n_runs) { failed = 1; continue; }
        run_first.push_back((uint16_t) ni);
        vis.assign((size_t) ni, 0);  // visiting order: longest item first (stable)
        {
          uint32_t cnt_len[17] = {0}, at[17] = {0};
          for (int i = 0; i < ni; ++i) ++cnt_len[((kitems[i] >> 16) & 15u) + 1u];
          uint32_t running = 0;
          for (int len = 16; len >= 1; --len) { at[len] = running; running += cnt_len[len]; }
          for (int i = 0; i < ni; ++i) vis[at[((kitems[i] >> 16) & 15u) + 1u]++] = (uint32_t) i;
        }
        // slices of 32 items, entries transposed (entry i of the slice's lane l at offset i * 32 + l), padded with the
        // zero slot (= local class id n_classes) up to the slice's padded length
        const int n_slices = (ni + 31) / 32;
        std::vector<uint32_t> slice_word((size_t) n_slices);
        int64_t n_words = 0;
        for (int sidx = 0; sidx < n_slices; ++sidx) {
          const int len0 = (int) ((kitems[vis[(size_t) sidx * 32]] >> 16) & 15u) + 1;
          const int L = kSliceLen[len0];
          slice_word[sidx] = (uint32_t) (n_words << 5) | (uint32_t) L;
          n_words += 32 * (int64_t) L;
        }
        if (n_words >= (int64_t(1) << 27)) { failed = 1; continue; }
        const int64_t off_pos = align_up(4 * (int64_t) n_slices, 16);
        const int64_t off_runkey = align_up(off_pos + 2 * (int64_t) ni, 16);
        const int64_t off_runfirst = align_up(off_runkey + 2 * (int64_t) nr, 16);
        const int64_t off_ents = align_up(off_runfirst + 2 * ((int64_t) nr + 1), 16);
        const int64_t b_bytes = align_up(off_ents + 2 * n_words, 16);
        std::vector<uint8_t>& out_bytes = bytes[u];
        out_bytes.assign((size_t) (a_bytes + b_bytes), 0);
        uint8_t* A = out_bytes.data();
        uint8_t* B = A + a_bytes;
        uint32_t* hdr = reinterpret_cast<uint32_t*>(A);
        hdr[GBRS_TH_CLASSES] = (uint32_t) t.n_classes;
        hdr[GBRS_TH_LOCI] = (uint32_t) t.n_loci;
        hdr[GBRS_TH_PLANES] = (uint32_t) t.n_planes;
        hdr[GBRS_TH_RUNS] = (uint32_t) nr;
        hdr[GBRS_TH_ITEMS] = (uint32_t) ni;
        hdr[GBRS_TH_SLICES] = (uint32_t) n_slices;
        hdr[GBRS_TH_A_BYTES] = (uint32_t) a_bytes;
        hdr[GBRS_TH_FULL] = full;
        hdr[GBRS_TH_PAIRS] = (uint32_t) t.n_pairs;
        hdr[GBRS_TH_ENTRIES] = (uint32_t) t.n_entries;
        hdr[GBRS_TH_B_BYTES] = (uint32_t) b_bytes;
        hdr[GBRS_TH_SELL_WORDS] = (uint32_t) n_words;
        std::memcpy(A + off_loci, loci.data(), 4 * loci.size());
        uint16_t* np16 = reinterpret_cast<uint16_t*>(A + off_nplane);
        for (int p = 0; p < t.n_planes; ++p) np16[p] = (uint16_t) nplane[p];
        double* cnt = reinterpret_cast<double*>(A + off_count);
        for (int j = 0; j < t.n_classes; ++j) cnt[j] = count[cls[j]];
        std::memcpy(A + off_pairs, pw.data(), 2 * pw.size());
        std::memcpy(B, slice_word.data(), 4 * slice_word.size());
        uint16_t* pos = reinterpret_cast<uint16_t*>(B + off_pos);
        for (int v = 0; v < ni; ++v) pos[v] = (uint16_t) vis[v];
        std::memcpy(B + off_runkey, run_key.data(), 2 * run_key.size());
        std::memcpy(B + off_runfirst, run_first.data(), 2 * run_first.size());
        uint16_t* ents = reinterpret_cast<uint16_t*>(B + off_ents);
        for (int sidx = 0; sidx < n_slices; ++sidx) {
          const int L = (int) (slice_word[sidx] & 31u);
          uint16_t* base = ents + (slice_word[sidx] >> 5);
          for (int lane = 0; lane < 32; ++lane) {
            const int v = sidx * 32 + lane;
            uint32_t st = 0, len = 0;
            if (v < ni) { st = kitems[vis[v]] & 0xFFFFu; len = ((kitems[vis[v]] >> 16) & 15u) + 1u; }
            for (int i = 0; i < L; ++i)
              base[i * 32 + lane] = (uint32_t) i < len ? (uint16_t) (keyed[st + i] & 0xFFFFu) : (uint16_t) t.n_classes;
          }
        }
        t.a_bytes = a_bytes;
        t.b_bytes = b_bytes;
        t.n_slices = n_slices;
        t.off_loci = (uint32_t) off_loci;
        t.off_slots = (uint32_t) off_slots;
        t.off_a[0] = (uint32_t) off_slots; t.off_a[1] = (uint32_t) off_nplane; t.off_a[2] = (uint32_t) off_count; t.off_a[3] = (uint32_t) off_pairs;
        t.off_b[0] = (uint32_t) off_pos; t.off_b[1] = (uint32_t) off_runkey; t.off_b[2] = (uint32_t) off_runfirst; t.off_b[3] = (uint32_t) off_ents;
      }
    }
    if (failed) { delete R; gbrs_set_error("gbrs_tiles_create: internal inconsistency while filling the tiles"); return GBRS_E_ARG; }

    // ---- 4. concatenate ------------------------------------------------------------------------------------------------
    int64_t blob_bytes = 0;
    for (TileStat& t : tiles) {
      t.blob_off = blob_bytes;
      blob_bytes = align_up(blob_bytes + t.a_bytes + t.b_bytes, 128);
      info.max_classes = std::max(info.max_classes, t.n_classes);
      info.max_loci = std::max(info.max_loci, t.n_loci);
      info.max_items = std::max(info.max_items, t.n_items);
      info.max_planes = std::max(info.max_planes, t.n_planes);
      info.max_part_a_bytes = std::max<int32_t>(info.max_part_a_bytes, (int32_t) t.a_bytes);
      info.max_part_b_bytes = std::max<int32_t>(info.max_part_b_bytes, (int32_t) t.b_bytes);
      info.n_entries += t.n_entries;
      info.n_items += t.n_items;
      info.n_pairs += t.n_pairs;
      info.n_classes += t.n_classes;
      info.n_slots += t.n_loci;
    }
    if (blob_bytes / 16 >= (int64_t(1) << 32)) { delete R; gbrs_set_error("gbrs_tiles_create: blob too large"); return GBRS_E_LIMIT; }
    info.n_tiles = n_tiles;
    info.blob_bytes = blob_bytes;
    info.item_len = ilen;
    R->blob.resize((size_t) blob_bytes);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t u = 0; u < n_tiles; ++u) {
      uint8_t* dst = R->blob.data() + tiles[u].blob_off;
      std::memcpy(dst, bytes[u].data(), bytes[u].size());
      const int64_t end = u + 1 < n_tiles ? tiles[u + 1].blob_off : blob_bytes;
      std::memset(dst + bytes[u].size(), 0, (size_t) (end - tiles[u].blob_off) - bytes[u].size());
      std::vector<uint8_t>().swap(bytes[u]);
    }

    // ---- 5. output slots, locus-major; locus descriptors ---------------------------------------------------------------
    std::vector<uint32_t> slot_ptr((size_t) T + 1, 0);
    for (const TileStat& t : tiles) {
      const uint32_t* loci = reinterpret_cast<const uint32_t*>(R->blob.data() + t.blob_off + t.off_loci);
      for (int l = 0; l < t.n_loci; ++l) ++slot_ptr[loci[l] + 1];
    }
    for (int t = 0; t < T; ++t) {
      info.max_slots_per_locus = std::max<int32_t>(info.max_slots_per_locus, (int32_t) slot_ptr[t + 1]);
      slot_ptr[t + 1] += slot_ptr[t];
    }
    {
      std::vector<uint32_t> cur(slot_ptr.begin(), slot_ptr.end() - 1);
      for (const TileStat& t : tiles) {
        const uint32_t* loci = reinterpret_cast<const uint32_t*>(R->blob.data() + t.blob_off + t.off_loci);
        uint32_t* slots = reinterpret_cast<uint32_t*>(R->blob.data() + t.blob_off + t.off_slots);
        for (int l = 0; l < t.n_loci; ++l) slots[l] = cur[loci[l]]++;
      }
    }
    {
      std::vector<uint32_t> lo((size_t) T);
      std::iota(lo.begin(), lo.end(), 0u);
      std::stable_sort(lo.begin(), lo.end(), [&](uint32_t x, uint32_t y) {
        return slot_ptr[x + 1] - slot_ptr[x] > slot_ptr[y + 1] - slot_ptr[y];
      });
      R->locus_desc.assign((size_t) T * 4, 0);
      for (int i = 0; i < T; ++i) {
        R->locus_desc[4 * (size_t) i + 0] = lo[i];
        R->locus_desc[4 * (size_t) i + 1] = slot_ptr[lo[i]];
        R->locus_desc[4 * (size_t) i + 2] = slot_ptr[lo[i] + 1];
        info.n_deep_loci += slot_ptr[lo[i] + 1] - slot_ptr[lo[i]] > GBRS_DEEP_LOCUS_ITEMS;
      }
    }
    // ---- 6. visiting order: costliest tile first (the work counter hands them out in this order) --------------------------
    {
      std::vector<uint32_t> vo((size_t) n_tiles);
      std::iota(vo.begin(), vo.end(), 0u);
      auto cost = [&](uint32_t u) { return (int64_t) tiles[u].n_pairs * 2 + tiles[u].n_entries + 4 * tiles[u].n_items + 8 * tiles[u].n_loci; };
      std::stable_sort(vo.begin(), vo.end(), [&](uint32_t x, uint32_t y) { return cost(x) > cost(y); });
      R->tile_desc.assign((size_t) std::max<int64_t>(n_tiles, 1) * GBRS_TD_WORDS, 0);
      for (int64_t i = 0; i < n_tiles; ++i) {
        const TileStat& t = tiles[vo[i]];
        uint32_t* dsc = R->tile_desc.data() + GBRS_TD_WORDS * (size_t) i;
        dsc[0] = (uint32_t) (t.blob_off / 16);
        dsc[1] = (uint32_t) t.n_classes | ((uint32_t) t.n_loci << 16);
        dsc[2] = (uint32_t) t.n_planes | ((uint32_t) t.n_runs << 16);
        dsc[3] = (uint32_t) t.n_items | ((uint32_t) t.n_slices << 16);
        dsc[4] = (uint32_t) t.a_bytes;
        dsc[5] = full;
        dsc[6] = vo[i];
        dsc[7] = (uint32_t) (t.a_bytes + t.b_bytes);
        for (int k = 0; k < 4; ++k) { dsc[8 + k] = t.off_a[k]; dsc[12 + k] = t.off_b[k]; }
      }
    }
    *out = R;
    return GBRS_OK;
  } catch (const std::bad_alloc&) {
    gbrs_set_error("gbrs_tiles_create: out of host memory");
    return GBRS_E_NOMEM;
  }
}

extern "C" int gbrs_tiles_get_info(gbrs_tiles_t t, gbrs_tiles_info* info) {
  if (!t || !info) { gbrs_set_error("gbrs_tiles_get_info: null argument"); return GBRS_E_ARG; }
  *info = t->info;
  return GBRS_OK;
}

extern "C" int gbrs_tiles_get_array(gbrs_tiles_t t, const char* name, const void** ptr, int64_t* bytes) {
  if (!t || !name || !ptr || !bytes) { gbrs_set_error("gbrs_tiles_get_array: null argument"); return GBRS_E_ARG; }
  const std::string s(name);
  if (s == "blob") { *ptr = t->blob.data(); *bytes = (int64_t) t->blob.size(); return GBRS_OK; }
  if (s == "tile_desc") { *ptr = t->tile_desc.data(); *bytes = (int64_t) (t->tile_desc.size() * 4); return GBRS_OK; }
  if (s == "locus_desc") { *ptr = t->locus_desc.data(); *bytes = (int64_t) (t->locus_desc.size() * 4); return GBRS_OK; }
  gbrs_set_error("gbrs_tiles_get_array: unknown array name '" + s + "'");
  return GBRS_E_ARG;
}

extern "C" int gbrs_tiles_free(gbrs_tiles_t t) {
  delete t;
  return GBRS_OK;
}
