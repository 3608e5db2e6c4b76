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
  int32_t n_classes = 0, n_loci = 0, n_pairs = 0, n_entries = 0, n_items = 0, n_planes = 0, n_runs = 0;
  int64_t a_bytes = 0, b_bytes = 0, blob_off = 0;
  uint32_t off_loci = 0, off_slots = 0, off_nplane = 0, off_count = 0, off_pairs = 0, off_ents = 0, off_pos = 0, off_runkey = 0,
           off_runfirst = 0;
};

// entries a pair word contributes to the tile's locus-major copy
inline int entries_of(uint32_t mask, uint32_t full) { return mask == full ? 1 : ((mask & 15u) != 0) + ((mask >> 4) != 0); }

}  // namespace

extern "C" int gbrs_tiles_create(gbrs_pack_t P, const gbrs_tiles_params* prm, gbrs_tiles_t* out) {
  if (!P || !out) { gbrs_set_error("gbrs_tiles_create: null argument"); return GBRS_E_ARG; }
  gbrs_tiles_params q{};
  if (prm) q = *prm;
  const int maxC = q.max_classes > 0 ? q.max_classes : 1024;
  const int maxL = q.max_loci > 0 ? q.max_loci : 32;
  const int maxP = q.max_pairs > 0 ? q.max_pairs : 3072;
  const int maxE = q.max_entries > 0 ? q.max_entries : 4608;
  const int maxI = q.max_items > 0 ? q.max_items : 1536;
  const int ilen = q.item_len > 0 ? q.item_len : 16;
  if (maxC > 2048 || maxL > 128 || maxE > 65535 || maxI > 65535 || ilen > 16 || maxP < maxL || maxE < 2 * maxL || maxI < 2 * maxL) {
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

    // ---- 3. blob geometry ------------------------------------------------------------------------------------------
    int64_t blob_bytes = 0;
    auto* R = new gbrs_tiles();
    gbrs_tiles_info& info = R->info;
    for (TileStat& t : tiles) {
      int64_t o = GBRS_TH_WORDS * 4;
      t.off_loci = (uint32_t) o;   o = align_up(o + 4 * (int64_t) t.n_loci, 16);
      t.off_slots = (uint32_t) o;  o = align_up(o + 4 * (int64_t) t.n_loci, 16);
      t.off_nplane = (uint32_t) o; o = align_up(o + 2 * (int64_t) t.n_planes, 16);
      t.off_count = (uint32_t) o;  o = align_up(o + 8 * (int64_t) t.n_classes, 16);
      // every plane is padded to a multiple of 4 words (one thread reads the words of 4 neighbouring classes at once)
      t.off_pairs = (uint32_t) o;  o = align_up(o + 2 * ((int64_t) t.n_pairs + 3 * (int64_t) t.n_planes), 16);
      t.a_bytes = o;
      int64_t ob = align_up(4 * (int64_t) t.n_items, 16);
      t.off_pos = (uint32_t) ob;      ob = align_up(ob + 2 * (int64_t) t.n_items, 16);
      t.off_runkey = (uint32_t) ob;   ob = align_up(ob + 2 * (int64_t) t.n_runs, 16);
      t.off_runfirst = (uint32_t) ob; ob = align_up(ob + 2 * ((int64_t) t.n_runs + 1), 16);
      t.off_ents = (uint32_t) ob;
      ob = align_up(ob + 2 * (int64_t) t.n_entries + 32, 16);  // + 32: the item loops may read up to 15 words past an item
      t.b_bytes = ob;
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
    par_fill(R->blob, (size_t) blob_bytes, (uint8_t) 0);

    // ---- 4. fill every tile (parallel); slots are assigned afterwards ---------------------------------------------------
    int failed = 0;
#pragma omp parallel
    {
      std::vector<uint32_t> loci, cls, keyed, kitems;
#pragma omp for schedule(dynamic, 8)
      for (int64_t u = 0; u < n_tiles; ++u) {
        const TileStat& t = tiles[u];
        uint8_t* A = R->blob.data() + t.blob_off;
        uint8_t* B = A + t.a_bytes;
        uint32_t* hdr = reinterpret_cast<uint32_t*>(A);
        hdr[GBRS_TH_CLASSES] = (uint32_t) t.n_classes;
        hdr[GBRS_TH_LOCI] = (uint32_t) t.n_loci;
        hdr[GBRS_TH_PLANES] = (uint32_t) t.n_planes;
        hdr[GBRS_TH_PAIRS] = (uint32_t) t.n_pairs;
        hdr[GBRS_TH_ENTRIES] = (uint32_t) t.n_entries;
        hdr[GBRS_TH_ITEMS] = (uint32_t) t.n_items;
        hdr[GBRS_TH_OFF_LOCI] = t.off_loci;
        hdr[GBRS_TH_OFF_SLOTS] = t.off_slots;
        hdr[GBRS_TH_OFF_NPLANE] = t.off_nplane;
        hdr[GBRS_TH_OFF_COUNT] = t.off_count;
        hdr[GBRS_TH_OFF_PAIRS] = t.off_pairs;
        hdr[GBRS_TH_A_BYTES] = (uint32_t) t.a_bytes;
        hdr[GBRS_TH_B_BYTES] = (uint32_t) t.b_bytes;
        hdr[GBRS_TH_OFF_ENTS] = t.off_ents;
        hdr[GBRS_TH_FLAGS] = full;  // the mask value that means "all haplotypes"
        hdr[GBRS_TH_OFF_POS] = t.off_pos;
        hdr[GBRS_TH_RUNS] = (uint32_t) t.n_runs;
        hdr[GBRS_TH_OFF_RUNKEY] = t.off_runkey;
        hdr[GBRS_TH_OFF_RUNFIRST] = t.off_runfirst;
        // locus list
        loci.clear();
        for (int64_t i = t.first; i < t.first + t.n_classes; ++i) {
          const uint32_t c = order[i];
          for (uint32_t p = rowptr[c]; p < rowptr[c + 1]; ++p) loci.push_back(pairs[p] & 0xFFFFFFu);
        }
        std::sort(loci.begin(), loci.end());
        loci.erase(std::unique(loci.begin(), loci.end()), loci.end());
        if ((int) loci.size() != t.n_loci) { failed = 1; continue; }
        std::memcpy(A + t.off_loci, loci.data(), 4 * loci.size());
        // classes by descending width (stable: smallest-locus order is kept inside a width)
        cls.assign(order.begin() + t.first, order.begin() + t.first + t.n_classes);
        std::stable_sort(cls.begin(), cls.end(), [&](uint32_t x, uint32_t y) {
          return rowptr[x + 1] - rowptr[x] > rowptr[y + 1] - rowptr[y];
        });
        uint16_t* nplane = reinterpret_cast<uint16_t*>(A + t.off_nplane);
        double* cnt = reinterpret_cast<double*>(A + t.off_count);
        uint16_t* pw = reinterpret_cast<uint16_t*>(A + t.off_pairs);
        for (int j = 0; j < t.n_classes; ++j) {
          const int k = (int) (rowptr[cls[j] + 1] - rowptr[cls[j]]);
          for (int p = 0; p < k; ++p) ++nplane[p];
          cnt[j] = count[cls[j]];
        }
        keyed.clear();
        {
          std::vector<uint32_t> plane_off((size_t) t.n_planes + 1, 0);
          for (int p = 0; p < t.n_planes; ++p) plane_off[p + 1] = plane_off[p] + ((uint32_t) nplane[p] + 3u) / 4u * 4u;
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
        }
        if ((int) keyed.size() != t.n_entries) { failed = 1; continue; }
        std::sort(keyed.begin(), keyed.end());
        // items in key order (scratch), runs per key
        uint16_t* ents = reinterpret_cast<uint16_t*>(B + t.off_ents);
        uint16_t* run_key = reinterpret_cast<uint16_t*>(B + t.off_runkey);
        uint16_t* run_first = reinterpret_cast<uint16_t*>(B + t.off_runfirst);
        kitems.clear();
        int nr = 0;
        for (size_t i = 0; i < keyed.size();) {
          const uint32_t key = keyed[i] >> 16;
          size_t j = i;
          while (j < keyed.size() && (keyed[j] >> 16) == key) ++j;
          if (nr < t.n_runs) { run_key[nr] = (uint16_t) key; run_first[nr] = (uint16_t) kitems.size(); }
          ++nr;
          for (size_t st = i; st < j; st += (size_t) ilen) {
            const uint32_t len = (uint32_t) std::min<size_t>((size_t) ilen, j - st);
            kitems.push_back((uint32_t) st | ((len - 1) << 16));
          }
          i = j;
        }
        const int ni = (int) kitems.size();
        if (ni != t.n_items || nr != t.n_runs) { failed = 1; continue; }
        run_first[nr] = (uint16_t) ni;
        for (size_t i = 0; i < keyed.size(); ++i) ents[i] = (uint16_t) (keyed[i] & 0xFFFFu);
        // visiting order of the items: longest first, so that the lanes of a warp walk items of (nearly) equal length
        uint32_t* items = reinterpret_cast<uint32_t*>(B);
        uint16_t* pos = reinterpret_cast<uint16_t*>(B + t.off_pos);
        {
          uint32_t cnt_len[17] = {0}, at[17] = {0};  // indexed by item length 1..16
          for (int i = 0; i < ni; ++i) ++cnt_len[((kitems[i] >> 16) & 15u) + 1u];
          uint32_t running = 0;
          for (int len = 16; len >= 1; --len) { at[len] = running; running += cnt_len[len]; }
          for (int i = 0; i < ni; ++i) {
            const uint32_t v = at[((kitems[i] >> 16) & 15u) + 1u]++;
            items[v] = kitems[i];
            pos[v] = (uint16_t) i;
          }
        }
      }
    }
    if (failed) { delete R; gbrs_set_error("gbrs_tiles_create: internal inconsistency while filling the tiles"); return GBRS_E_ARG; }

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
      }
    }
    // ---- 6. visiting order: costliest tile first (the work counter hands them out in this order) --------------------------
    {
      std::vector<uint32_t> vo((size_t) n_tiles);
      std::iota(vo.begin(), vo.end(), 0u);
      auto cost = [&](uint32_t u) { return (int64_t) tiles[u].n_pairs * 2 + tiles[u].n_entries + 4 * tiles[u].n_items + 8 * tiles[u].n_loci; };
      std::stable_sort(vo.begin(), vo.end(), [&](uint32_t x, uint32_t y) { return cost(x) > cost(y); });
      R->tile_desc.assign((size_t) std::max<int64_t>(n_tiles, 1) * 4, 0);
      for (int64_t i = 0; i < n_tiles; ++i) {
        const TileStat& t = tiles[vo[i]];
        R->tile_desc[4 * (size_t) i + 0] = (uint32_t) (t.blob_off / 16);
        R->tile_desc[4 * (size_t) i + 1] = (uint32_t) t.a_bytes;
        R->tile_desc[4 * (size_t) i + 2] = (uint32_t) t.b_bytes;
        R->tile_desc[4 * (size_t) i + 3] = vo[i];
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
