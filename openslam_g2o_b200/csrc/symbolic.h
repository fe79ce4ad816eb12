// symbolic.h - host-side symbolic analysis for the supernodal left-looking block Cholesky.
//
// Replaces (one-time, per structure) what the reference does in
// LinearSolverCSparse::computeSymbolicDecomposition (solvers/csparse/linear_solver_csparse.h:246-300):
// block AMD -> permuted pattern -> elimination tree -> column counts (nnz(L)), and adds what a GPU
// numeric phase needs and CSparse's scalar up-looking factorisation does not: supernodes (fundamental,
// relaxed amalgamation, width cap), per-supernode row structures, left-looking update lists with
// relative indices, and a task/level schedule (small subtrees run inside one CTA, the top of the tree is
// level-scheduled).
#pragma once
#include <cstdint>
#include <vector>

namespace g2o_b200 {

struct SymbolicOptions {
  int max_panel_cols_scalar = 72;   // supernodes wider than this are split into a chain of panels (<= 72 scalars
                                    // and <= 12 blocks: the panel factorisation maps one warp per block column)
  double subtree_work_fraction = 1.0 / 1024;  // subtree tasks: at most this share of the total work
  double subtree_min_flops = 5.0e5;           // ... but never split below what one CTA does in ~10 us
  double subtree_max_flops = 2.0e7;           // ... and never more than this per task (large graphs; measured: 90k-pose sphere 76.8 -> 59.9 ms)
  bool relax = true;
  double relax_frac = 0.25;         // relaxed amalgamation: accepted share of explicit zeros up to one panel width
  bool sort_items_by_level = true;  // tile work items ordered by the task level of their source supernode
  bool groups_asap = true;          // a split-K group is listed one level above its latest source, not at its destination's level
  bool split_late_items = false;    // the items fed by the level just below the destination form their own group(s):
                                    // measured 3-4 % slower on sphere2500 and on the 250k-pose sphere (more split tiles)
  int group_slack = 0;              // ... plus this many levels (see symbolic.cpp; measured: 1 or 2 levels of slack are 5 % slower)
  int wide_group_items = 32;        // ... with wide tiles (unless group_items is set explicitly)
  int group_items = 16;             // split-K: work items per group task (4 / 8 / 16 measured: 16 best with level-sorted items)
  // 0 (default): the reference's ordering - block AMD, bit-exact with cs_amd.  k > 0: nested dissection with 2^k parts
  // on top of it (nested_dissection.cpp): separators first cut band-like systems, whose AMD elimination tree is one
  // long chain, into independent subtrees; AMD orders the parts and the separators
  int nd_levels = 0;
  int nd_min_part = 24;             // parts smaller than this many blocks are not cut further
  // destination tiles of the update plan: 0 = 48 x 48 (latency-bound factorisations: Venice, sphere2500), 1 = 96 x 72
  // with the products on the FP64 tensor path (DMMA), -1 = by the factorisation's flop count (d = 6 only)
  int wide_tiles = -1;
  double wide_min_flops = 2.0e10;
  // set when the caller already knows the ordering (tests); empty = run block AMD
  std::vector<int> given_perm;
  // tail chain (chol.cu: chol_chain_kernel): the last stretch of the elimination tree - the path from some supernode up
  // to the root - is factored by ONE CTA that keeps the whole frontal matrix in registers and never signals through
  // HBM between links.  Only block dimension 6, fronts of at most chain_max_rows block rows.
  bool chain = true;
  int chain_max_rows = 29;   // = kChR of chol_chain.cuh
  int chain_min_links = 3;
  size_t chain_smem_budget = 190 * 1024;  // staged panel + re-index buffer of the widest link (+ ~31 KB of fixed buffers) must fit
};

struct SymbolicFactor {
  int nb = 0;  // block columns
  int d = 0;   // block dimension (3 or 6)
  std::vector<int> perm, pinv;  // perm[new] = old block index
  // block-level elimination tree / exact column counts of the permuted matrix
  std::vector<int> parent, colcount;
  int64_t scalar_lnz = 0;  // == css::lnz of the reference for the same ordering
  // scatter plan of the input blocks (input order) into the panels of L
  std::vector<int64_t> a_dst;     // offset of element (0,0) of the destination block
  std::vector<int32_t> a_ld;      // leading dimension of the destination panel
  std::vector<uint8_t> a_trans;   // 1: store the transposed block
  std::vector<int64_t> diag_dst;  // per permuted block column: offset of its diagonal block (for lambda)
  std::vector<int32_t> diag_ld;
  // supernodes (ascending column order)
  int nsn = 0;
  std::vector<int> sn_col0, sn_ncol, sn_nrow;  // block units; nrow includes the ncol diagonal rows
  std::vector<int> sn_rowptr, sn_rows;         // block rows, ascending; first ncol entries = own columns
  std::vector<int64_t> sn_lptr;                // offset of the panel in L ((nrow*d) x (ncol*d), col-major)
  std::vector<int> sn_parent, col2sn;
  // left-looking updates: target J receives from K rows [p0,p1) (the rows of K inside J's columns)
  std::vector<int> upd_ptr;                    // nsn+1
  std::vector<int> upd_k, upd_p0, upd_p1;
  std::vector<int64_t> upd_relptr;             // into rel
  std::vector<int> rel;                        // for p in [p0,nrow_K): local block row in J
  // schedule: tasks = supernode sequences processed by one CTA; levels of independent tasks
  std::vector<int> task_ptr, task_sn;          // task t runs task_sn[task_ptr[t] .. task_ptr[t+1])
  std::vector<int> level_ptr;                  // tasks sorted by level; level l = tasks [level_ptr[l], level_ptr[l+1])
  int nlevels = 0;
  int64_t factor_doubles = 0;
  double flops = 0;  // factorisation flops of the stored (relaxed) structure
  double subtree_flops = 0;  // ... of which inside SUBTREE tasks (one CTA each)
  int max_nrow = 0, max_ncol = 0;

  // ---- numeric plan of the GPU kernels (chol.cu)
  // destination tiles: a panel is cut into tile_blocks x tile_blocks block tiles (48 x 48 scalars); only tiles that
  // touch the lower triangle exist.  Each tile owns the list of update pieces (work items) that land in it.
  int tile_blocks = 0;                             // block rows of a tile
  int tile_blocks_c = 0;                           // block columns of a tile (= tile_blocks unless wide)
  bool wide = false;                               // 96 x 72 tiles + FP64 tensor-path products (SymbolicOptions::wide_tiles)
  std::vector<int> sn_tile_ptr;                    // nsn+1
  std::vector<int> tile_sn, tile_r0, tile_c0;      // supernode, first local block row / block column
  std::vector<int> tile_work_ptr;                  // ntiles+1
  std::vector<int> work_u, work_a0, work_a1, work_b0, work_b1;  // update index; row / column ranges relative to p0
  // the same work items flattened for the device (one level of indirection instead of four)
  std::vector<int64_t> work_koff, work_reloff;     // offset of row p0 of the updating panel in L; offset into rel
  std::vector<int> work_mk, work_nk;               // leading dimension / #columns of the updating panel (scalars)
  // row chunks of the panel factorisation: every chunk CTA factors the diagonal block and solves its block rows
  int chunk_blocks = 0;
  std::vector<int> sn_chunk_ptr;                   // nsn+1
  std::vector<int> chunk_sn, chunk_b0, chunk_nb;   // supernode, first local block row (>= ncol), #block rows
  // level plan (intermediate of the dataflow task list): kind 0 = subtree tasks (one CTA does update + factor of
  // every supernode of the task), 1 = split (group tasks per tile, chunk tasks per panel)
  std::vector<int> level_kind;
  std::vector<int> level_chunk_ptr, level_chunks;
  // split-K groups of the split levels: a tile with many work items is cut into groups of <= group_items items;
  // each group is one CTA.  slot < 0: the tile has a single group and subtracts from the panel directly; otherwise
  // the group writes its partial 48x48 sum to scratch slot `slot` (numbered globally: levels overlap in the dataflow
  // kernel) and the tile's reduce task adds the slots in order.
  int group_items = 0;
  std::vector<int> level_group_ptr, group_tile, group_w0, group_w1, group_slot;   // per level: its groups
  std::vector<int> level_rtile_ptr, rtile_tile, rtile_slot0, rtile_nslots;        // per level: tiles needing a reduce
  int max_group_slots = 0;
  // forward solve: supernode J leaves L21*y_J (its contribution to every ancestor) at sn_cptr[J]
  std::vector<int64_t> sn_cptr;                    // nsn+1, scalar units
  // ... and every scalar row of the permuted system lists the contribution entries it has to subtract, in
  // ascending (supernode, row) order (fixed summation order)
  std::vector<int> fwd_ptr, fwd_src;               // n+1 ; indices into the contribution array
  // ---- dataflow schedule (chol.cu: one persistent kernel per factorisation / per backward sweep).  The numeric work
  // is cut into tasks listed in level-major order; a CTA takes the next task from a global counter and spins on the
  // completion counters of what the task consumes.  Every dependency points to an earlier entry of the list, so
  // the earliest unfinished task is always held by a running CTA (no deadlock, whatever the grid size).
  //   kind 0 SUBTREE(task)  update + factor of every supernode of a small subtree, sequentially in one CTA
  //   kind 1 GROUP(group)   one split-K group of a destination tile: waits for each source supernode in list order;
  //                         the group of a split tile that finishes last adds the partial sums in group order
  //                         and subtracts once (rtile_* arrays)
  //   kind 3 CHUNK(chunk)   waits for every update of its supernode, factors diagonal block + its row chunk
  std::vector<int> flow_kind, flow_arg;
  std::vector<int> work_ksn;                       // source supernode of a work item
  std::vector<int> sn_nupd;                        // update completions (direct groups + rtiles) a supernode waits for
  std::vector<int> sn_nchunk;                      // chunks of a supernode: it is ready when all of them are stored
  std::vector<int> group_rtile;                    // per group: its rtile (-1: subtracts directly)
  std::vector<int> task_parent;                    // per task: the task holding the parent of its root supernode
  // inverses of the triangular diagonal blocks (for the solves)
  std::vector<int64_t> sn_dinvptr;                 // nsn+1
  int64_t dinv_doubles = 0;

  // ---- tail chain (see SymbolicOptions::chain).  chain_sn lists the supernodes of the path bottom -> top (the last one is
  // the root); they appear in no CHUNK task and no work item has one of them as its SOURCE: what a chain supernode
  // receives from the supernodes below the chain still arrives through GROUP tasks in its HBM panel, what it receives
  // from earlier links stays in the chain CTA's registers (multifrontal: the update matrix of link j, re-indexed by
  // chain_map, is the start of link j+1's front).
  std::vector<int> chain_sn;
  std::vector<char> sn_on_chain;                   // nsn
  std::vector<int> chain_mapptr, chain_map;        // per link: for every block row BELOW its diagonal block, the local row
                                                   // in the next link's front (empty for the last link)
  std::vector<unsigned> chain_new_rows;            // per link: bit r = local row r enters the front at this link (zero it)
  std::vector<int> chain_colptr;                   // nlinks+1: scalar columns before link j (into chain_fwd_ptr)
  std::vector<int> chain_fwd_ptr, chain_fwd_src;   // per scalar column of the chain: contribution entries of sources
                                                   // below the chain (forward substitution)
  std::vector<char> task_on_chain;                 // per task: 1 = its (single) supernode is a chain link
  size_t chain_stage_doubles = 0, chain_remap_blocks = 0;  // shared-memory needs of the widest link
  double chain_flops = 0;
};

// nested_dissection.cpp
std::vector<int> nested_dissection_order(int nb, const int* colptr, const int* rowidx, int levels, int min_part);

// colptr/rowidx: upper block pattern (rows <= col, ascending, diagonal present) in input block order.
SymbolicFactor analyze(int nb, int d, const int* colptr, const int* rowidx, const SymbolicOptions& opt);

}  // namespace g2o_b200
