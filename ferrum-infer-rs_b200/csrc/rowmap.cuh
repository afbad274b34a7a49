// rowmap.cuh -- device-side resolution of "row blocks" (see RowMap in common.cuh).
#pragma once
#include "common.cuh"

namespace fb2 {

// The routing arrays are read with ld.global.cg (L2-coherent), never through the non-coherent
// read-only path: the fused MoE kernel produces them earlier IN THE SAME LAUNCH.
// Number of row blocks that exist for this launch (device side; ROUTED reads the padded
// length the align kernel produced, so no host sync is needed).
__device__ __forceinline__ int rowmap_num_blocks(const RowMap &rm, int nb) {
  if (rm.mode == ROWS_DENSE) return (rm.m + nb - 1) / nb;
  if (rm.mode == ROWS_DISPATCH) return rm.num_dispatch;
  return __ldcg(rm.num_post_padded) / rm.block;
}

// Expert of block b (0 for dense).
__device__ __forceinline__ int rowmap_block_expert(const RowMap &rm, const DispatchTable *tbl,
                                                   int b) {
  if (rm.mode == ROWS_DENSE) return 0;
  if (rm.mode == ROWS_DISPATCH) return tbl->e[b].expert;
  return __ldcg(rm.block_experts + b);
}

// Rows of block b: slot i in [0, nb).  Returns false for padding slots.
//   a_row: row of the activation matrix to read;  c_row: row of the output to write.
__device__ __forceinline__ bool rowmap_resolve(const RowMap &rm, const DispatchTable *tbl, int nb,
                                               int b, int i, int &a_row, int &c_row) {
  if (rm.mode == ROWS_DENSE) {
    int r = b * nb + i;
    a_row = c_row = r;
    return r < rm.m;
  }
  if (rm.mode == ROWS_DISPATCH) {
    const DispatchEntry &d = tbl->e[b];
    a_row = d.in_row + i;
    c_row = d.out_row + i;
    return i < d.m;
  }
  if (i >= rm.block) return false;
  int id = __ldcg(rm.sorted_ids + (size_t)b * rm.block + i);
  a_row = id / rm.top_k;   // marlin_template.h contract: A row = id / top_k
  c_row = id;              //                            C row = id
  return id < rm.num_valid;
}

}  // namespace fb2
