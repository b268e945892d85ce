// Host-only check of the stream-K work split of mu_gemm_sm100.cuh: the segment numbering the kernel's roles follow
// (pieces -> cta_range -> decode) against the slot lists reduce_slots_of_tile() hands to the reduce kernels.
// usage: worksplit_check num_tiles kb_per_tile piece_len grid   -> prints "OK <max_segs>" or an error
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#include "../../alpine_b200/csrc/mu_gemm_sm100.cuh"

using namespace alpine;

int main(int argc, char** argv) {
  if (argc != 5) return 2;
  WorkSpace ws;
  ws.num_tiles = atoi(argv[1]);
  ws.kb_per_tile = atoi(argv[2]);
  ws.piece_len = atoi(argv[3]);
  ws.pieces = (ws.kb_per_tile + ws.piece_len - 1) / ws.piece_len;
  const int grid = atoi(argv[4]);
  const int max_segs = max_segments_per_cta(ws, grid);
  // what the kernel does: per CTA, per piece, contiguous runs inside a tile, numbered from 0
  std::map<std::pair<int, int>, std::vector<std::pair<int, long long>>> by_run;  // (piece, tile) -> (slot, k-blocks)
  long long covered = 0;
  for (int cta = 0; cta < grid; ++cta) {
    int seg = 0;
    for (int pc = 0; pc < ws.pieces; ++pc) {
      long long b, e;
      ws.cta_range(pc, grid, cta, b, e);
      for (long long pos = b; pos < e; ++seg) {
        int run, tile, kb0, len;
        ws.decode(pos, run, tile, kb0, len);
        if (len > e - pos) len = static_cast<int>(e - pos);
        if (run != pc * ws.num_tiles + tile || kb0 < pc * ws.piece_len || kb0 + len > ws.kb_per_tile) {
          printf("bad decode at %lld\n", pos);
          return 1;
        }
        if (seg >= max_segs) {
          printf("cta %d exceeds max_segs %d\n", cta, max_segs);
          return 1;
        }
        by_run[{pc, tile}].push_back({cta * max_segs + seg, len});
        covered += len;
        pos += len;
      }
    }
  }
  if (covered != ws.total()) {
    printf("covered %lld of %lld\n", covered, ws.total());
    return 1;
  }
  std::set<int> used;
  for (int tile = 0; tile < ws.num_tiles; ++tile) {
    std::vector<int> slots;
    reduce_slots_of_tile(ws, grid, max_segs, tile, &slots);
    std::vector<int> expect;
    long long kb = 0;
    for (int pc = 0; pc < ws.pieces; ++pc)
      for (auto& s : by_run[{pc, tile}]) {
        expect.push_back(s.first);
        kb += s.second;
      }
    if (kb != ws.kb_per_tile) {
      printf("tile %d reduces %lld of %d k-blocks\n", tile, kb, ws.kb_per_tile);
      return 1;
    }
    if (slots != expect) {
      printf("tile %d: slot list differs (%zu vs %zu)\n", tile, slots.size(), expect.size());
      return 1;
    }
    for (int s : slots)
      if (!used.insert(s).second) {
        printf("slot %d used twice\n", s);
        return 1;
      }
  }
  printf("OK %d\n", max_segs);
  return 0;
}
