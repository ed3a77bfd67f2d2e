// CUDA kernels of the assembly hot path (sm_100a, FP64):
//   K1  cell kernel (pe_cell.cuh)  +  K2 scatter-add into the CSR values  +  K3 per-cell Dirichlet
//   elimination, fused.  Replaces the cell loop of assemble_system and
//   AffineConstraints::distribute_local_to_global (PoroElasBR1WG_new.cc:791-1170,
//   PoroElasCGQ1_WG.cc:1165-1725).
#include "pe_internal.h"
#include "pe_kernels.h"

#include <cuda_pipeline.h>

#include <algorithm>
#include <cstdlib>

namespace pe
{
  // ---------------------------------------------------------------------------------------------
  // setup kernel: for every assembled cell a and local pair (i,j): the position of the entry inside
  // the CSR row of dof i (7 bits) and whether a is the FIRST cell, in launch order, that
  // contributes to that entry (bit 7).  First writers store, later ones add: no memset of the
  // matrix, no atomics, and a fixed summation order (bitwise reproducible assembly).
  // The assembled cells are ordered colour-major, so "first" = smallest asm index among the cells
  // that contain both dofs.  row_cells = cells adjacent to every owned row (asm indices).
  // ---------------------------------------------------------------------------------------------
  __global__ void
  k_scatter_offsets(const int n_loc, const int64_t n_asm, const int64_t L, const int32_t *__restrict__ cdofs,
                    const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                    const int64_t *__restrict__ rc_ptr, const int32_t *__restrict__ rc_idx,
                    const uint16_t *__restrict__ slot /* [n_loc*n_loc] */, uint8_t *__restrict__ off /* [4*words][n_asm] */,
                    uint32_t *__restrict__ rhs_first, int *__restrict__ err,
                    const uint8_t *__restrict__ nzs /* [n_loc*n_loc] non-zero slot or 255 */,
                    uint8_t *__restrict__ gsrc /* [nnz][4] or null */, int *__restrict__ gather_bad)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_asm * n_loc)
      return;
    const int64_t c = t % n_asm;
    const int     i = (int)(t / n_asm);
    const int32_t r = cdofs[(size_t)i * n_asm + c];
    if (r >= L)
      {
        for (int j = 0; j < n_loc; ++j)
          {
            const int sl = slot[i * n_loc + j];
            off[((size_t)(sl >> 2) * n_asm + c) * 4 + (sl & 3)] = 0;
          }
        return;
      }
    const int64_t b = rowptr[r], e = rowptr[r + 1];
    if (e - b > 128)
      {
        atomicExch(err, 2);
        return;
      }
    const int64_t cb = rc_ptr[r], ce = rc_ptr[r + 1];
    // rhs: first cell touching the row
    {
      int64_t first = c;
      for (int64_t k = cb; k < ce; ++k)
        first = min(first, (int64_t)rc_idx[k]);
      if (first == c)
        atomicOr(rhs_first + c, 1u << i);
    }
    for (int j = 0; j < n_loc; ++j)
      {
        const int32_t col = cdofs[(size_t)j * n_asm + c];
        // linear search: with local column numbering (multi-GPU) a row is not sorted
        int64_t lo = b;
        while (lo < e && colidx[lo] != col)
          ++lo;
        if (lo >= e)
          {
            atomicExch(err, 1);
            lo = b;
          }
        // smallest asm index among the cells of this row that also contain `col`
        int64_t first = c;
        for (int64_t k = cb; k < ce; ++k)
          {
            const int64_t o = rc_idx[k];
            if (o >= first)
              continue;
            bool has = false;
            for (int jj = 0; jj < n_loc; ++jj)
              has |= cdofs[(size_t)jj * n_asm + o] == col;
            if (has)
              first = o;
          }
        const int sl = slot[i * n_loc + j];
        off[((size_t)(sl >> 2) * n_asm + c) * 4 + (sl & 3)] = (uint8_t)((lo - b) | (first == c ? 128 : 0));
        // gather table: CSR entry (r, col) receives non-zero slot nzs(i,j) of the s-th cell adjacent to row r
        if (gsrc)
          {
            int sidx = -1;
            for (int64_t k = cb; k < ce; ++k)
              if (rc_idx[k] == c)
                sidx = (int)(k - cb);
            if (ce - cb > 4 || sidx < 0)
              atomicExch(gather_bad, 1); // a row with more than four adjacent cells: no gather assembly
            else if (nzs[i * n_loc + j] != 255)
              gsrc[(size_t)lo * 4 + sidx] = nzs[i * n_loc + j];
          }
      }
  }

  // ---------------------------------------------------------------------------------------------
  // emit functors
  // ---------------------------------------------------------------------------------------------
  template <int EL, int NL, bool MAT, bool ATOMIC>
  struct ColorEmit // first writer stores; later colours add with RED (no return value -> no load latency
                   // in the dependency chain; uncontended and ordered by the colour launches -> deterministic)
  {
    double         *values, *rhsp;
    const uint32_t *soff; // shared memory: word w of this thread at soff[w * 128]; words relative to the sweep
    int             word0; // first word of the sweep
    uint32_t        rhs_first;
    int64_t        rs[NL]; // CSR row start of local dof i, -1 if the row is not owned
    int32_t        row[NL];
    __device__ __forceinline__ void
    mat(const int i, const int j, const double v)
    {
      if (!MAT || rs[i] < 0)
        return;
      if (ATOMIC && v == 0.)
        return; // pre-zeroed matrix: structural zeros need no traffic
      const int      sl = SlotMap<EL>::slot(i, j);
      const uint32_t o  = (soff[((sl >> 2) - word0) * 128] >> ((sl & 3) * 8)) & 255u;
      double        *p  = values + rs[i] + (o & 127u);
      if (!ATOMIC && (o & 128u))
        *p = v;
      else if (v != 0.)
        atomicAdd(p, v); // RED: fire-and-forget
    }
    __device__ __forceinline__ void
    rhs(const int i, const double v)
    {
      if (rs[i] < 0 || rhsp == nullptr) // rhsp == nullptr: matrix-only pass (solver operator A0)
        return;
      if (!ATOMIC && ((rhs_first >> i) & 1u))
        rhsp[row[i]] = v;
      else if (v != 0.)
        atomicAdd(rhsp + row[i], v);
    }
  };

  template <int NL>
  struct LocalEmit // into thread-local arrays (boundary cells)
  {
    double *A, *b;
    __device__ __forceinline__ void
    mat(const int i, const int j, const double v)
    {
      A[i * NL + j] = v;
    }
    __device__ __forceinline__ void
    rhs(const int i, const double v)
    {
      b[i] = v;
    }
  };

  template <int NL>
  struct SoaEmit // dense local matrices, entry-major [NL*NL][n] (BASELINE cfg 5)
  {
    double *out, *rhs_out;
    size_t  n;
    __device__ __forceinline__ void
    mat(const int i, const int j, const double v)
    {
      out[(size_t)(i * NL + j) * n] = v;
    }
    __device__ __forceinline__ void
    rhs(const int i, const double v)
    {
      if (rhs_out)
        rhs_out[(size_t)i * n] = v;
    }
  };

  template <int EL>
  __device__ __forceinline__ void
  load_cell(const AsmArgs &a, const int64_t c, double vx[4], double vy[4], int32_t idx[Elem<EL>::NL],
            double old[Elem<EL>::NL])
  {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      {
        const uint32_t vi = a.cv[(size_t)v * a.n_asm + c];
        vx[v]             = __ldg(a.vx + vi);
        vy[v]             = __ldg(a.vy + vi);
      }
#pragma unroll
    for (int i = 0; i < Elem<EL>::NL; ++i)
      {
        idx[i] = a.cdofs[(size_t)i * a.n_asm + c];
        old[i] = a.sol ? __ldg(a.sol + idx[i]) : 0.;
      }
  }

  // ---------------------------------------------------------------------------------------------
  // interior cells (no boundary face, no constrained dof): K1 + K2, one launch per colour and sweep
  // ---------------------------------------------------------------------------------------------
  template <int EL, int SW, bool MAT, bool ATOMIC>
  __device__ __forceinline__ void
  assemble_interior_body(const AsmArgs &a, const DevParams &P)
  {
    constexpr int NL = Elem<EL>::NL;
    const int64_t c  = a.a_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.a_end)
      return;
    if (a.flags[c] != 0u)
      return; // handled by k_assemble_boundary
    double  vx[4], vy[4], old[NL];
    int32_t idx[NL];
    load_cell<EL>(a, c, vx, vy, idx, old);
    Geom G;
    G.set(vx, vy);
    // stage this sweep's scatter-offset words through shared memory: independent coalesced loads
    constexpr int SWI = SW == 1 ? 0 : SW == 2 ? 1 : 2;
    constexpr int W0 = SlotMap<EL>::word_begin(SWI), W1 = SlotMap<EL>::word_begin(SWI + 1);
    __shared__ uint32_t soff[(MAT ? W1 - W0 : 1) * 128];
    if (MAT)
      {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.off) + (size_t)W0 * a.n_asm + c;
#pragma unroll
        for (int w = 0; w < W1 - W0; ++w)
          soff[w * 128 + threadIdx.x] = __ldg(src + (size_t)w * a.n_asm);
      }
    ColorEmit<EL, NL, MAT, ATOMIC> E;
    E.values    = a.values;
    E.rhsp      = a.rhs;
    E.soff      = soff + threadIdx.x;
    E.word0     = W0;
    E.rhs_first = a.rhs_first[c];
#pragma unroll
    for (int i = 0; i < NL; ++i)
      {
        E.row[i] = idx[i];
        E.rs[i]  = idx[i] < a.L ? __ldg(a.rowptr + idx[i]) : -1;
      }
    cell_kernel<EL, SW>(G, a.kcell ? a.kcell[c] : a.k_scalar, P, old, E);
  }
  template <int EL, int SW, bool MAT, bool ATOMIC>
  __global__ void __launch_bounds__(128)
  k_assemble_interior(const AsmArgs a, const DevParams P)
  {
    assemble_interior_body<EL, SW, MAT, ATOMIC>(a, P);
  }
  // the three sweeps of one chunk of cells in ONE launch (blockIdx.y = sweep)
  template <int EL>
  __global__ void __launch_bounds__(128)
  k_assemble_interior3(const AsmArgs a, const DevParams P)
  {
    if (blockIdx.y == 0)
      assemble_interior_body<EL, 1, true, true>(a, P);
    else if (blockIdx.y == 1)
      assemble_interior_body<EL, 2, true, true>(a, P);
    else
      assemble_interior_body<EL, 4, true, true>(a, P);
  }
  // zero up to three ranges of the CSR values (the rows first touched by a chunk of cells)
  __global__ void
  k_zero_ranges(double *__restrict__ v, const int64_t b0, const int64_t e0, const int64_t b1, const int64_t e1,
                const int64_t b2, const int64_t e2)
  {
    const int64_t n0 = e0 - b0, n1 = e1 - b1, n2 = e2 - b2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1 + n2; i += (int64_t)gridDim.x * blockDim.x)
      {
        const int64_t k = i < n0 ? b0 + i : (i < n0 + n1 ? b1 + (i - n0) : b2 + (i - n0 - n1));
        v[k]            = 0.;
      }
  }

  // ---------------------------------------------------------------------------------------------
  // Gather assembly (opt-in alternative, PE_ASM_GATHER=1, meshes whose rows touch <= 4 cells): no atomics on the
  // bulk of the entries.  MEASURED SLOWER than the RED scatter on B200 (N = 1024: 4.15 ms against 2.23 ms): the
  // staged contributions are written coalesced but read back one 32-byte sector per 8-byte entry, i.e. the same
  // 193 M L2 transactions as the REDs plus the staging traffic.
  //   phase A  per chunk of Morton-consecutive cells: the cell kernels stage every structurally non-zero local
  //            entry at T[nz][cell in chunk] with plain coalesced stores (the same access pattern as the dense
  //            local-matrix kernel, 61 % of the HBM roof).  Rows first touched by an EARLIER chunk (2 % of the
  //            cells of a 32 Ki chunk sit on its low boundary) are updated with RED as before.
  //   phase B  one warp per row first touched by the chunk: lane e sums, for CSR entry e of the row, the staged
  //            contributions of the <= 4 adjacent cells of the chunk (one table word per entry: four slot bytes)
  //            and stores the value -- which also replaces the memset of the matrix.  Cells of LATER chunks add
  //            their share with RED when their chunk runs (stream order), boundary cells at the very end.
  // ---------------------------------------------------------------------------------------------
  struct GatherArgs
  {
    double        *T;        // [NZ][chunk_stride]
    int64_t        chunk_stride;
    int64_t        row_begin[3], row_end[3]; // local rows first touched by this chunk, per block (u | p_int | p_face)
    const uint8_t *gsrc;     // [nnz][4]
    const int32_t *rowadj;   // [L][4] asm indices of the cells adjacent to a row, -1 padded
    // tiled phase B (k_assemble_gatherB2): tiles of kTileCells consecutive cells of the chunk
    const int64_t *tile_rows;      // [n_tiles][6] local rows first touched by the tile, per block
    const int32_t *tile_ghost_ptr; // [n_tiles + 1]
    const int32_t *tile_ghost;     // asm indices of the cells OUTSIDE the tile (same chunk) adjacent to its rows
    int64_t        tile_first;     // index of the first tile of this chunk
  };

  template <int EL, int NL>
  struct TileEmit
  {
    double        *T, *values, *rhsp;
    const uint8_t *off;
    int64_t        n_asm, c;
    uint32_t       own; // bit i: local dof i is a row first touched by this chunk
    int64_t        rs[NL];
    int32_t        row[NL];
    __device__ __forceinline__ void
    rhs(const int i, const double v)
    {
      if (rs[i] < 0 || rhsp == nullptr)
        return;
      if (v != 0.)
        atomicAdd(rhsp + row[i], v);
    }
  };
  // T is indexed [nz * stride + cell]; T points at this cell's column
  template <int EL, int NL>
  struct TileEmitS : TileEmit<EL, NL>
  {
    int64_t stride;
    __device__ __forceinline__ void
    mat(const int i, const int j, const double v)
    {
      const int nz = nzslot<EL>(i, j);
      if (nz < 0 || this->rs[i] < 0)
        return;
      if ((this->own >> i) & 1u)
        this->T[(size_t)nz * stride] = v;
      else if (v != 0.)
        {
          const int      sl = SlotMap<EL>::slot(i, j);
          const uint32_t o  = this->off[((size_t)(sl >> 2) * this->n_asm + this->c) * 4 + (sl & 3)];
          atomicAdd(this->values + this->rs[i] + (o & 127u), v);
        }
    }
  };

  template <int EL, int SW>
  __device__ __forceinline__ void
  assemble_gather_body(const AsmArgs &a, const DevParams &P, const GatherArgs &g)
  {
    constexpr int NL = Elem<EL>::NL;
    const int64_t t  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t c  = a.a_begin + t;
    if (c >= a.a_end)
      return;
    double *Tc = g.T + t;
    if (a.flags[c] != 0u)
      {
        // boundary cells are assembled by k_assemble_boundary at the end: their staged contribution is zero
        if (SW == 1)
          for (int z = 0; z < NzCount<EL>::value; ++z)
            Tc[(size_t)z * g.chunk_stride] = 0.;
        return;
      }
    double  vx[4], vy[4], old[NL];
    int32_t idx[NL];
    load_cell<EL>(a, c, vx, vy, idx, old);
    Geom G;
    G.set(vx, vy);
    TileEmitS<EL, NL> E;
    E.T      = Tc;
    E.stride = g.chunk_stride;
    E.values = a.values;
    E.rhsp   = a.rhs;
    E.off    = a.off;
    E.n_asm  = a.n_asm;
    E.c      = c;
    E.own    = 0u;
#pragma unroll
    for (int i = 0; i < NL; ++i)
      {
        E.row[i] = idx[i];
        E.rs[i]  = idx[i] < a.L ? __ldg(a.rowptr + idx[i]) : -1;
        const int blk = SlotMap<EL>::kind(i) <= 1 ? 0 : (SlotMap<EL>::kind(i) == 3 ? 1 : 2);
        if (idx[i] >= g.row_begin[blk] && idx[i] < g.row_end[blk])
          E.own |= 1u << i;
      }
    cell_kernel<EL, SW>(G, a.kcell ? a.kcell[c] : a.k_scalar, P, old, E);
  }
  template <int EL>
  __global__ void __launch_bounds__(128)
  k_assemble_gatherA(const AsmArgs a, const DevParams P, const GatherArgs g)
  {
    if (blockIdx.y == 0)
      assemble_gather_body<EL, 1>(a, P, g);
    else if (blockIdx.y == 1)
      assemble_gather_body<EL, 2>(a, P, g);
    else
      assemble_gather_body<EL, 4>(a, P, g);
  }
  // phase B: warp per row
  __global__ void __launch_bounds__(256)
  k_assemble_gatherB(const AsmArgs a, const GatherArgs g)
  {
    const int64_t w    = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int     lane = threadIdx.x & 31;
    const int64_t n0 = g.row_end[0] - g.row_begin[0], n1 = g.row_end[1] - g.row_begin[1], n2 = g.row_end[2] - g.row_begin[2];
    if (w >= n0 + n1 + n2)
      return;
    const int64_t r = w < n0 ? g.row_begin[0] + w : (w < n0 + n1 ? g.row_begin[1] + (w - n0) : g.row_begin[2] + (w - n0 - n1));
    // cells of this chunk adjacent to the row (column index inside T), -1: none / a later chunk
    int64_t cell[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      {
        const int32_t ac = __ldg(g.rowadj + 4 * r + q);
        cell[q]          = (ac >= a.a_begin && ac < a.a_end) ? ac - a.a_begin : -1;
      }
    const int64_t b = __ldg(a.rowptr + r), e = __ldg(a.rowptr + r + 1);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(g.gsrc);
    for (int64_t k = b + lane; k < e; k += 32)
      {
        const uint32_t wq = __ldg(src + k);
        double         v  = 0.;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          {
            const uint32_t nz = (wq >> (8 * q)) & 255u;
            if (nz != 255u && cell[q] >= 0)
              v += g.T[(size_t)nz * g.chunk_stride + cell[q]];
          }
        a.values[k] = v;
      }
  }

  // phase B, tiled: one CTA per tile of kTileCells Morton-consecutive cells.  The staged entries of the tile's cells
  // (coalesced: 32 consecutive cells per slot) and of the few cells outside the tile that touch its rows are loaded
  // into shared memory once; every CSR entry of the rows first touched by the tile is then summed from shared memory
  // (<= 4 contributions, one table word per entry) and stored -- each matrix value is written exactly once, coalesced,
  // with no atomics and in a fixed order (bitwise reproducible).  Replaces the warp-per-row phase B whose read-back
  // cost one 32-byte L2 sector per 8-byte contribution.
  template <int NZ>
  __global__ void __launch_bounds__(kTileThreads)
  k_assemble_gatherB2(const AsmArgs a, const GatherArgs g)
  {
    constexpr int            NT = kTileThreads, NW = NT / 32;
    extern __shared__ double S[]; // [kTileCells + kTileGhostMax][NZ]
    constexpr int            kRowsMax = 320; // rows of one tile (all three blocks); longer tiles run in batches
    __shared__ int32_t       ghost[kTileGhostMax];
    __shared__ int64_t       rs_sh[3][kRowsMax + 1];
    __shared__ int8_t        cs_sh[3][kRowsMax][4];
    const int64_t tile = g.tile_first + blockIdx.x;
    const int64_t c0   = a.a_begin + (int64_t)blockIdx.x * kTileCells;
    const int     nown = (int)min((int64_t)kTileCells, a.a_end - c0);
    const int     g0 = g.tile_ghost_ptr[tile], ng = g.tile_ghost_ptr[tile + 1] - g0;
    const int     lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- all global -> shared copies of the tile are issued asynchronously (cp.async, 8 bytes each) before anything
    // waits: own cells warp <-> slot, lane <-> cell (256 B coalesced per slot); cells outside the tile lane <-> slot
    const int64_t t0 = c0 - a.a_begin; // column of the first own cell in T
    for (int nz = warp; nz < NZ; nz += NW)
      if (lane < nown)
        __pipeline_memcpy_async(&S[lane * NZ + nz], &g.T[(size_t)nz * g.chunk_stride + t0 + lane], sizeof(double));
    for (int q = warp; q < ng; q += NW)
      {
        const int64_t tc = (int64_t)__ldg(g.tile_ghost + g0 + q) - a.a_begin;
        for (int nz = lane; nz < NZ; nz += 32)
          __pipeline_memcpy_async(&S[(kTileCells + q) * NZ + nz], &g.T[(size_t)nz * g.chunk_stride + tc], sizeof(double));
      }
    __pipeline_commit();
    if (threadIdx.x < ng)
      ghost[threadIdx.x] = g.tile_ghost[g0 + threadIdx.x];
    const int64_t  *tr  = g.tile_rows + 6 * tile;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(g.gsrc);
    int64_t         done[3] = {tr[0], tr[2], tr[4]};
    bool            first   = true;
    while (done[0] < tr[1] || done[1] < tr[3] || done[2] < tr[5])
      {
        // row starts of the three blocks (contiguous CSR ranges)
        int nrow[3];
#pragma unroll
        for (int blk = 0; blk < 3; ++blk)
          {
            nrow[blk] = (int)min((int64_t)kRowsMax, tr[2 * blk + 1] - done[blk]);
            for (int t = threadIdx.x; t <= nrow[blk]; t += NT)
              rs_sh[blk][t] = __ldg(a.rowptr + done[blk] + t);
          }
        __syncthreads(); // ghost[] visible
#pragma unroll
        for (int blk = 0; blk < 3; ++blk)
          for (int t = threadIdx.x; t < nrow[blk] * 4; t += NT)
            {
              const int32_t ac = __ldg(g.rowadj + 4 * done[blk] + t);
              int           sl = -1;
              if (ac >= c0 && ac < c0 + nown)
                sl = (int)(ac - c0);
              else if (ac >= a.a_begin && ac < a.a_end)
                for (int z = 0; z < ng; ++z)
                  if (ghost[z] == ac)
                    sl = kTileCells + z;
              cs_sh[blk][t >> 2][t & 3] = (int8_t)sl; // -1: no cell / a cell of a later chunk (RED when its chunk runs)
            }
        if (first)
          __pipeline_wait_prior(0);
        first = false;
        __syncthreads();
        // one thread per CSR entry; the table words of four entries are loaded before any is used
#pragma unroll
        for (int blk = 0; blk < 3; ++blk)
          {
            const int64_t k0 = rs_sh[blk][0], k1 = rs_sh[blk][nrow[blk]];
            for (int64_t kb = k0 + threadIdx.x; kb < k1; kb += 4 * NT)
              {
                uint32_t wq[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  wq[u] = kb + u * NT < k1 ? __ldg(src + kb + u * NT) : 0xffffffffu;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  {
                    const int64_t k = kb + u * NT;
                    if (k >= k1)
                      break;
                    int lo = 0, hi = nrow[blk];
                    while (hi - lo > 1)
                      {
                        const int mid = (lo + hi) >> 1;
                        if (rs_sh[blk][mid] <= k)
                          lo = mid;
                        else
                          hi = mid;
                      }
                    double v = 0.;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                      {
                        const uint32_t nz = (wq[u] >> (8 * q)) & 255u;
                        const int      sl = cs_sh[blk][lo][q];
                        if (nz != 255u && sl >= 0)
                          v += S[sl * NZ + nz];
                      }
                    a.values[k] = v;
                  }
              }
            done[blk] += nrow[blk];
          }
        __syncthreads();
      }
  }

  // ---------------------------------------------------------------------------------------------
  // boundary cells: K1 into thread-local storage, then AffineConstraints::distribute_local_to_global
  // with use_inhomogeneities_for_rhs = true (deal.II 9.1 affine_constraints.templates.h; SURVEY
  // 8(a) row A8):
  //   unconstrained row r:  Mat(r,c) += A(r,c) for unconstrained c;  rhs(r) += b(r) - sum_c A(r,c) g_c
  //   constrained dof c  :  d = |A(c,c)| != 0 ? |A(c,c)| : mean_i |A(i,i)| ;  Mat(c,c) += d ; rhs(c) += d g_c
  // Entries dropped by the elimination still have to be written (as zeros) by their first writer.
  // ---------------------------------------------------------------------------------------------
  template <int EL>
  __global__ void __launch_bounds__(64)
  k_assemble_boundary(const AsmArgs a, const DevParams P)
  {
    constexpr int NL = Elem<EL>::NL;
    const int64_t t  = a.bnd_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.bnd_end)
      return;
    const int64_t c = a.bnd_list[t];
    double        vx[4], vy[4], old[NL];
    int32_t       idx[NL];
    load_cell<EL>(a, c, vx, vy, idx, old);
    Geom G;
    G.set(vx, vy);
    double        A[NL * NL], b[NL], g[NL];
    bool          con[NL];
    LocalEmit<NL> E{A, b};
    cell_kernel<EL, 7>(G, a.kcell ? a.kcell[c] : a.k_scalar, P, old, E);
    cell_boundary<EL>(G, vx, vy, a.flags[c], P, b, con, g);

    double avg = 0.;
    for (int i = 0; i < NL; ++i)
      avg += fabs(A[i * NL + i]);
    avg /= (double)NL;
    const uint32_t rf = a.rhs_first[c];
    for (int i = 0; i < NL; ++i)
      {
        if (idx[i] >= a.L)
          continue; // row owned by another rank
        const int64_t rs = a.rowptr[idx[i]];
        auto          o  = [&](const int j) -> uint32_t {
          const int sl = SlotMap<EL>::slot(i, j);
          return a.off[((size_t)(sl >> 2) * a.n_asm + c) * 4 + (sl & 3)];
        };
        double         val;
        if (con[i])
          {
            const double d = fabs(A[i * NL + i]) != 0. ? fabs(A[i * NL + i]) : avg;
            val            = d * g[i];
            if (a.do_matrix)
              for (int j = 0; j < NL; ++j)
                {
                  const uint32_t ob = o(j);
                  double        *p  = a.values + rs + (ob & 127u);
                  const double   v  = j == i ? d : 0.;
                  if (!a.atomic_mode && (ob & 128u))
                    *p = v;
                  else if (v != 0.)
                    atomicAdd(p, v);
                }
          }
        else
          {
            val = b[i];
            for (int j = 0; j < NL; ++j)
              {
                double v = A[i * NL + j];
                if (con[j])
                  {
                    if (g[j] != 0.)
                      val -= v * g[j];
                    v = 0.;
                  }
                if (a.do_matrix)
                  {
                    const uint32_t ob = o(j);
                    double        *p  = a.values + rs + (ob & 127u);
                    if (!a.atomic_mode && (ob & 128u))
                      *p = v;
                    else if (v != 0.)
                      atomicAdd(p, v);
                  }
              }
          }
        if (a.rhs == nullptr)
          continue;
        if (!a.atomic_mode && ((rf >> i) & 1u))
          a.rhs[idx[i]] = val;
        else if (val != 0.)
          atomicAdd(a.rhs + idx[i], val);
      }
  }

  // ---------------------------------------------------------------------------------------------
  // local matrices only (BASELINE cfg 5 and the parity hook pe_local_matrices)
  // ---------------------------------------------------------------------------------------------
  template <int EL, int SW>
  __global__ void __launch_bounds__(128)
  k_local_matrices(const AsmArgs a, const DevParams P, const int64_t first, const int64_t n, double *out,
                   double *rhs_out)
  {
    constexpr int NL = Elem<EL>::NL;
    const int64_t t  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
      return;
    const int64_t c = first + t;
    double        vx[4], vy[4], old[NL];
#pragma unroll
    for (int v = 0; v < 4; ++v)
      {
        const uint32_t vi = a.cv[(size_t)v * a.n_asm + c];
        vx[v]             = __ldg(a.vx + vi);
        vy[v]             = __ldg(a.vy + vi);
      }
#pragma unroll
    for (int i = 0; i < NL; ++i)
      old[i] = (a.sol && a.cdofs) ? __ldg(a.sol + a.cdofs[(size_t)i * a.n_asm + c]) : 0.;
    Geom G;
    G.set(vx, vy);
    SoaEmit<NL> E{out + t, rhs_out ? rhs_out + t : nullptr, (size_t)n};
    cell_kernel<EL, SW>(G, a.kcell ? a.kcell[c] : a.k_scalar, P, old, E);
  }

  // ---------------------------------------------------------------------------------------------
  // postprocess_errors (PoroElasBR1WG_new.cc:1251-1349, PoroElasCGQ1_WG.cc:2229-2357): per-step squared
  // errors of the device solution against the manufactured solution, QGauss(3) on every owned cell:
  //   e[0] = ||p_int - p||^2   e[1] = ||u_h - u||^2   e[2] = ||grad u_h - grad u||^2
  // (integrate_difference + compute_global_error = sqrt of the sum over cells of the cell integrals)
  // ---------------------------------------------------------------------------------------------
  template <int EL>
  __global__ void __launch_bounds__(128)
  k_step_errors(const AsmArgs a, const DevParams P, const uint8_t *__restrict__ is_ghost, double *__restrict__ e)
  {
    constexpr int NL = Elem<EL>::NL, NS = Elem<EL>::NS;
    const int64_t c  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double        ep = 0., eu = 0., eg = 0.;
    if (c < a.n_asm && !(is_ghost && is_ghost[c]))
      {
        double  vx[4], vy[4], sol[NL];
        int32_t idx[NL];
        load_cell<EL>(a, c, vx, vy, idx, sol);
        Geom G;
        G.set(vx, vy);
        const double ph = sol[Elem<EL>::PINT];
#pragma unroll 1
        for (int qi = 0; qi < 9; ++qi)
          {
            QPoint<NS> q;
            eval_qpoint<NS>(G, qi, q);
            double s[NS];
            shape_values<NS>(q.xi, q.eta, s);
            const double idet = q.wh / q.jxw; // 1/det
            double       ux = 0., uy = 0., gxx = 0., gxy = 0., gyx = 0., gyy = 0.;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              {
                ux += sol[2 * k] * s[k];
                uy += sol[2 * k + 1] * s[k];
                gxx += sol[2 * k] * q.nx[k];
                gxy += sol[2 * k] * q.ny[k];
                gyx += sol[2 * k + 1] * q.nx[k];
                gyy += sol[2 * k + 1] * q.ny[k];
              }
            if (EL == 0)
              {
#pragma unroll
                for (int f = 0; f < 4; ++f)
                  {
                    const double cf = sol[(8 + 2 * f) % NL];
                    if (f < 2)
                      {
                        ux += cf * s[(4 + f) % NS];
                        gxx += cf * q.nx[(4 + f) % NS];
                        gxy += cf * q.ny[(4 + f) % NS];
                      }
                    else
                      {
                        uy += cf * s[(4 + f) % NS];
                        gyx += cf * q.nx[(4 + f) % NS];
                        gyy += cf * q.ny[(4 + f) % NS];
                      }
                  }
              }
            gxx *= idet;
            gxy *= idet;
            gyx *= idet;
            gyy *= idet;
            double px, py, ex = 0., ey = 0., exx = 0., exy = 0., pe = 0.;
            G.point(q.xi, q.eta, px, py);
            if (P.case_id == 1)
              {
                double sx, cx, sy, cy;
                sincospi(px, &sx, &cx);
                sincospi(py, &sy, &cy);
                ex  = P.u_coef * sx * cy;
                ey  = P.u_coef * cx * sy;
                exx = P.u_coef * M_PI * cx * cy;  // d u_x / dx = d u_y / dy   (BR1new:278-283)
                exy = -P.u_coef * M_PI * sx * sy; // d u_x / dy = d u_y / dx
                pe  = P.p_coef * cx * cy;
              }
            ep += (ph - pe) * (ph - pe) * q.jxw;
            eu += ((ux - ex) * (ux - ex) + (uy - ey) * (uy - ey)) * q.jxw;
            eg += ((gxx - exx) * (gxx - exx) + (gxy - exy) * (gxy - exy) + (gyx - exy) * (gyx - exy) +
                   (gyy - exx) * (gyy - exx)) *
                  q.jxw;
          }
      }
    __shared__ double sh[3][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      {
        ep += __shfl_xor_sync(0xffffffffu, ep, o);
        eu += __shfl_xor_sync(0xffffffffu, eu, o);
        eg += __shfl_xor_sync(0xffffffffu, eg, o);
      }
    if ((threadIdx.x & 31) == 0)
      {
        sh[0][threadIdx.x >> 5] = ep;
        sh[1][threadIdx.x >> 5] = eu;
        sh[2][threadIdx.x >> 5] = eg;
      }
    __syncthreads();
    if (threadIdx.x < 3)
      atomicAdd(e + threadIdx.x, sh[threadIdx.x][0] + sh[threadIdx.x][1] + sh[threadIdx.x][2] + sh[threadIdx.x][3]);
  }

  cudaError_t
  launch_step_errors(const int element, const AsmArgs &a, const DevParams &P, const uint8_t *is_ghost, double *e,
                     cudaStream_t s)
  {
    cudaMemsetAsync(e, 0, 3 * sizeof(double), s);
    if (a.n_asm == 0)
      return cudaSuccess;
    const unsigned g = (unsigned)((a.n_asm + 127) / 128);
    if (element == 0)
      k_step_errors<0><<<g, 128, 0, s>>>(a, P, is_ghost, e);
    else
      k_step_errors<1><<<g, 128, 0, s>>>(a, P, is_ghost, e);
    return cudaGetLastError();
  }

  // ---------------------------------------------------------------------------------------------
  // postprocess() (PoroElasCGQ1_WG.cc:1744-1966; PoroElasBR1WG_new.cc:1189-1246 and
  // postprocess_darcy_velocity 1352-1606), one thread per assembled cell, results of owned cells only:
  //   div[c]  = (1/|E|) sum_q div u_h JxW          mag[c] = sqrt((1/|E|) sum_q |u_h|^2 JxW)      QGauss(3)
  //   beta[c] = RT0 coefficients of the Darcy velocity  -K grad_w p_h = -sum_i p_i C_ij D_kj w_k
  //   vel[c]  = that velocity at the cell midpoint (QMidpoint)
  // Lean form of the reference's loops (checked against the literal oracle): on any quadrilateral the RT0
  // weak-gradient right-hand side F is constant, and for K = k I the matrix D = M^-1 E is k I, so
  //   beta = -k M^-1 g,   g = (p_int - p_f0, p_f1 - p_int, p_int - p_f2, p_f3 - p_int).
  // ---------------------------------------------------------------------------------------------
  template <int EL>
  __global__ void __launch_bounds__(128)
  k_postprocess(const AsmArgs a, const int32_t *__restrict__ out_slot, double *__restrict__ div_out,
                double *__restrict__ mag_out, double *__restrict__ beta_out, double *__restrict__ vel_out)
  {
    constexpr int NL = Elem<EL>::NL, NS = Elem<EL>::NS;
    const int64_t c  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_asm)
      return;
    const int32_t o = out_slot[c];
    if (o < 0)
      return; // ghost cell
    double  vx[4], vy[4], sol[NL];
    int32_t idx[NL];
    load_cell<EL>(a, c, vx, vy, idx, sol);
    Geom G;
    G.set(vx, vy);
    double area = 0., dsum = 0., m2 = 0., M[10];
#pragma unroll
    for (int k = 0; k < 10; ++k)
      M[k] = 0.;
#pragma unroll 1
    for (int qi = 0; qi < 9; ++qi)
      {
        QPoint<NS> q;
        eval_qpoint<NS>(G, qi, q);
        double s[NS];
        shape_values<NS>(q.xi, q.eta, s);
        double ux = 0., uy = 0., dv = 0.;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          {
            ux += sol[2 * k] * s[k];
            uy += sol[2 * k + 1] * s[k];
            dv += sol[2 * k] * q.nx[k] + sol[2 * k + 1] * q.ny[k];
          }
        if (EL == 0)
          {
#pragma unroll
            for (int f = 0; f < 4; ++f)
              {
                const double cf = sol[(8 + 2 * f) % NL];
                if (f < 2)
                  {
                    ux += cf * s[(4 + f) % NS];
                    dv += cf * q.nx[(4 + f) % NS];
                  }
                else
                  {
                    uy += cf * s[(4 + f) % NS];
                    dv += cf * q.ny[(4 + f) % NS];
                  }
              }
          }
        area += q.jxw;
        dsum += dv * q.wh; // div = dv / det, JxW = wh det
        m2 += (ux * ux + uy * uy) * q.jxw;
        // RT0: w_k = J what_k / det, what = ((1-xi),0), (xi,0), (0,1-eta), (0,eta);  M += w_i . w_j JxW
        const double c0x = q.j00, c0y = q.j10, c1x = q.j01, c1y = q.j11; // columns of J
        const double a00 = (c0x * c0x + c0y * c0y) * q.r, a01 = (c0x * c1x + c0y * c1y) * q.r,
                     a11 = (c1x * c1x + c1y * c1y) * q.r;
        const double mxi = 1. - q.xi, meta = 1. - q.eta;
        M[sym4(0, 0)] += mxi * mxi * a00;
        M[sym4(0, 1)] += mxi * q.xi * a00;
        M[sym4(1, 1)] += q.xi * q.xi * a00;
        M[sym4(0, 2)] += mxi * meta * a01;
        M[sym4(0, 3)] += mxi * q.eta * a01;
        M[sym4(1, 2)] += q.xi * meta * a01;
        M[sym4(1, 3)] += q.xi * q.eta * a01;
        M[sym4(2, 2)] += meta * meta * a11;
        M[sym4(2, 3)] += meta * q.eta * a11;
        M[sym4(3, 3)] += q.eta * q.eta * a11;
      }
    double Mi[10];
    spd4_inverse(M, Mi);
    const double pint = sol[Elem<EL>::PINT];
    const double g[4] = {pint - sol[Elem<EL>::pface(0)], sol[Elem<EL>::pface(1)] - pint, pint - sol[Elem<EL>::pface(2)],
                         sol[Elem<EL>::pface(3)] - pint};
    const double kperm = a.kcell ? a.kcell[c] : a.k_scalar;
    double       beta[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      {
        double t = 0.;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          t += Mi[sym4(k, j)] * g[j];
        beta[k] = -kperm * t;
      }
    if (div_out)
      div_out[o] = dsum / area;
    if (mag_out)
      mag_out[o] = sqrt(m2 / area);
    if (beta_out)
      {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          beta_out[(size_t)4 * o + k] = beta[k];
      }
    if (vel_out)
      {
        double j00, j01, j10, j11;
        G.jac(.5, .5, j00, j01, j10, j11);
        const double idet = 1. / (j00 * j11 - j01 * j10);
        const double hx = .5 * (beta[0] + beta[1]), hy = .5 * (beta[2] + beta[3]); // sum_k beta_k what_k(1/2,1/2)
        vel_out[(size_t)2 * o]     = (j00 * hx + j01 * hy) * idet;
        vel_out[(size_t)2 * o + 1] = (j10 * hx + j11 * hy) * idet;
      }
  }

  cudaError_t
  launch_postprocess(const int element, const AsmArgs &a, const int32_t *out_slot, double *div_out, double *mag_out,
                     double *beta_out, double *vel_out, cudaStream_t s)
  {
    if (a.n_asm == 0)
      return cudaSuccess;
    const unsigned g = (unsigned)((a.n_asm + 127) / 128);
    if (element == 0)
      k_postprocess<0><<<g, 128, 0, s>>>(a, out_slot, div_out, mag_out, beta_out, vel_out);
    else
      k_postprocess<1><<<g, 128, 0, s>>>(a, out_slot, div_out, mag_out, beta_out, vel_out);
    return cudaGetLastError();
  }

  // K = exp(sigma Z), Z by Box-Muller from the counter generator keyed on (seed, global cell id)
  __device__ __forceinline__ uint64_t
  d_splitmix64(uint64_t z)
  {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  __device__ __forceinline__ double
  d_u01(uint64_t seed, uint64_t id, uint64_t k)
  {
    const uint64_t r = d_splitmix64(d_splitmix64(seed ^ (id * 0xD1B54A32D192ED03ull)) + k);
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __global__ void
  k_lognormal(const int64_t n, const int64_t *__restrict__ cell_ids, const int64_t first, const double sigma,
              const uint64_t seed, double *__restrict__ k)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
      return;
    const uint64_t id = (uint64_t)(cell_ids ? cell_ids[t] : first + t);
    const double   u1 = d_u01(seed, id, 0), u2 = d_u01(seed, id, 1);
    const double   z  = sqrt(-2. * log(u1)) * cospi(2. * u2);
    k[t]              = exp(sigma * z);
  }

  __global__ void
  k_sum(const double *__restrict__ x, const size_t n, double *__restrict__ out)
  {
    double s = 0.;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      s += x[i];
    for (int o = 16; o > 0; o >>= 1)
      s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double sh[32];
    if ((threadIdx.x & 31) == 0)
      sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
      {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1)
          s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0)
          atomicAdd(out, s);
      }
  }

  // ---------------------------------------------------------------------------------------------
  // launchers
  // ---------------------------------------------------------------------------------------------
  static inline unsigned
  grid_for(const int64_t n, const int block)
  {
    return (unsigned)((n + block - 1) / block);
  }

  cudaError_t
  launch_scatter_offsets(const int n_loc, const int64_t n_asm, const int64_t L, const int32_t *cdofs,
                         const int64_t *rowptr, const int32_t *colidx, const int64_t *rc_ptr, const int32_t *rc_idx,
                         uint8_t *off, uint32_t *rhs_first, int *err, uint8_t *gsrc, int *gather_bad, cudaStream_t s)
  {
    const int64_t n = n_asm * n_loc;
    if (n == 0)
      return cudaSuccess;
    uint16_t slot_h[21 * 21];
    for (int i = 0; i < n_loc; ++i)
      for (int j = 0; j < n_loc; ++j)
        slot_h[i * n_loc + j] = (uint16_t)(n_loc == 17 ? SlotMap<0>::slot(i, j) : (n_loc == 13 ? SlotMap<1>::slot(i, j) : i * n_loc + j));
    uint8_t nz_h[21 * 21];
    for (int i = 0; i < n_loc; ++i)
      for (int j = 0; j < n_loc; ++j)
        {
          const int z         = n_loc == 17 ? nzslot<0>(i, j) : (n_loc == 13 ? nzslot<1>(i, j) : -1);
          nz_h[i * n_loc + j] = z < 0 ? 255 : (uint8_t)z;
        }
    uint16_t *slot_d = nullptr;
    uint8_t  *nz_d   = nullptr;
    cudaError_t     e      = cudaMalloc((void **)&slot_d, sizeof(uint16_t) * n_loc * n_loc);
    if (e != cudaSuccess)
      return e;
    e = cudaMalloc((void **)&nz_d, n_loc * n_loc);
    if (e != cudaSuccess)
      return e;
    cudaMemcpyAsync(slot_d, slot_h, sizeof(uint16_t) * n_loc * n_loc, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(nz_d, nz_h, n_loc * n_loc, cudaMemcpyHostToDevice, s);
    cudaMemsetAsync(rhs_first, 0, sizeof(uint32_t) * (size_t)n_asm, s);
    k_scatter_offsets<<<grid_for(n, 256), 256, 0, s>>>(n_loc, n_asm, L, cdofs, rowptr, colidx, rc_ptr, rc_idx, slot_d, off,
                                                        rhs_first, err, nz_d, (n_loc == 17 || n_loc == 13) ? gsrc : nullptr,
                                                        gather_bad);
    e = cudaGetLastError();
    cudaStreamSynchronize(s);
    cudaFree(slot_d);
    cudaFree(nz_d);
    return e;
  }

  int
  scatter_table_words(const int n_loc)
  {
    return n_loc == 17 ? SlotMap<0>::word_begin(3) : (n_loc == 13 ? SlotMap<1>::word_begin(3) : (n_loc * n_loc + 3) / 4);
  }

  template <int EL>
  static void
  launch_color(AsmArgs a, const DevParams &P, cudaStream_t s, int *n_launches)
  {
    const int64_t n = a.a_end - a.a_begin;
    if (n > 0)
      {
        const unsigned g = grid_for(n, 128);
        if (a.do_matrix && a.atomic_mode)
          {
            k_assemble_interior<EL, 1, true, true><<<g, 128, 0, s>>>(a, P);
            k_assemble_interior<EL, 2, true, true><<<g, 128, 0, s>>>(a, P);
            k_assemble_interior<EL, 4, true, true><<<g, 128, 0, s>>>(a, P);
            *n_launches += 3;
          }
        else if (a.do_matrix)
          {
            k_assemble_interior<EL, 1, true, false><<<g, 128, 0, s>>>(a, P);
            k_assemble_interior<EL, 2, true, false><<<g, 128, 0, s>>>(a, P);
            k_assemble_interior<EL, 4, true, false><<<g, 128, 0, s>>>(a, P);
            *n_launches += 3;
          }
        else if (a.atomic_mode)
          {
            k_assemble_interior<EL, 4, false, true><<<g, 128, 0, s>>>(a, P);
            *n_launches += 1;
          }
        else
          {
            k_assemble_interior<EL, 4, false, false><<<g, 128, 0, s>>>(a, P);
            *n_launches += 1;
          }
      }
    if (a.bnd_end > a.bnd_begin)
      {
        k_assemble_boundary<EL><<<grid_for(a.bnd_end - a.bnd_begin, 64), 64, 0, s>>>(a, P);
        ++*n_launches;
      }
  }

  // atomics mode: chunks of cells in deal.II (Morton) order; the CSR rows first touched by a chunk are
  // zeroed right before the chunk is assembled, so that the zeros are still in L2 when the REDs
  // arrive (no DRAM fetch) and every value line is written back once.
  cudaError_t
  launch_assemble_chunked(const int element, const AsmArgs &a0, const DevParams &P, const int n_chunks,
                          const int64_t *chunk_cell_begin, const int64_t *chunk_zero /* [n_chunks][6] */,
                          cudaStream_t s, int *n_launches)
  {
    for (int ch = 0; ch < n_chunks; ++ch)
      {
        AsmArgs a = a0;
        a.a_begin = chunk_cell_begin[ch];
        a.a_end   = chunk_cell_begin[ch + 1];
        const int64_t *z = chunk_zero + 6 * ch;
        const int64_t  nz = (z[1] - z[0]) + (z[3] - z[2]) + (z[5] - z[4]);
        if (nz > 0)
          {
            k_zero_ranges<<<(unsigned)std::min<int64_t>((nz + 255) / 256, 148 * 16), 256, 0, s>>>(a.values, z[0], z[1], z[2],
                                                                                                  z[3], z[4], z[5]);
            ++*n_launches;
          }
        const int64_t n = a.a_end - a.a_begin;
        // (the Y sweep in a launch of its own, so that X and Q run at 3 CTAs per SM, measured slower: 2.41 ms
        //  against 2.23 ms at N = 1024)
        if (n > 0)
          {
            const dim3 g(grid_for(n, 128), 3);
            if (element == 0)
              k_assemble_interior3<0><<<g, 128, 0, s>>>(a, P);
            else
              k_assemble_interior3<1><<<g, 128, 0, s>>>(a, P);
            ++*n_launches;
          }
      }
    // boundary cells last (their rows are all zeroed by now)
    AsmArgs a   = a0;
    a.bnd_begin = 0;
    a.bnd_end   = a0.n_bnd;
    if (a.n_bnd > 0)
      {
        if (element == 0)
          k_assemble_boundary<0><<<grid_for(a.n_bnd, 64), 64, 0, s>>>(a, P);
        else
          k_assemble_boundary<1><<<grid_for(a.n_bnd, 64), 64, 0, s>>>(a, P);
        ++*n_launches;
      }
    return cudaGetLastError();
  }

  cudaError_t
  launch_assemble_gather(const int element, const AsmArgs &a0, const DevParams &P, const int n_chunks,
                         const int64_t *chunk_cell_begin, const int64_t *chunk_rows /* [n_chunks][6] */, double *T,
                         const int64_t chunk_stride, const uint8_t *gsrc, const int32_t *rowadj, const GatherTiles *tiles,
                         cudaStream_t s, int *n_launches)
  {
    for (int ch = 0; ch < n_chunks; ++ch)
      {
        AsmArgs a = a0;
        a.a_begin = chunk_cell_begin[ch];
        a.a_end   = chunk_cell_begin[ch + 1];
        GatherArgs g;
        g.T            = T;
        g.chunk_stride = chunk_stride;
        g.gsrc         = gsrc;
        g.rowadj       = rowadj;
        int64_t n_rows = 0;
        for (int b = 0; b < 3; ++b)
          {
            g.row_begin[b] = chunk_rows[6 * ch + 2 * b];
            g.row_end[b]   = chunk_rows[6 * ch + 2 * b + 1];
            n_rows += g.row_end[b] - g.row_begin[b];
          }
        const int64_t n = a.a_end - a.a_begin;
        if (n > 0)
          {
            const dim3 grid(grid_for(n, 128), 3);
            if (element == 0)
              k_assemble_gatherA<0><<<grid, 128, 0, s>>>(a, P, g);
            else
              k_assemble_gatherA<1><<<grid, 128, 0, s>>>(a, P, g);
            ++*n_launches;
          }
        if (n_rows > 0 && tiles && tiles->tile_rows)
          {
            g.tile_rows      = tiles->tile_rows;
            g.tile_ghost_ptr = tiles->tile_ghost_ptr;
            g.tile_ghost     = tiles->tile_ghost;
            g.tile_first     = a.a_begin / kTileCells;
            const unsigned nt = (unsigned)((n + kTileCells - 1) / kTileCells);
            if (element == 0)
              {
                constexpr int    NZ   = NzCount<0>::value;
                constexpr size_t smem = sizeof(double) * (size_t)(kTileCells + kTileGhostMax) * NZ;
                static bool      attr = false;
                if (!attr)
                  {
                    cudaFuncSetAttribute(k_assemble_gatherB2<NZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    attr = true;
                  }
                k_assemble_gatherB2<NZ><<<nt, kTileThreads, smem, s>>>(a, g);
              }
            else
              {
                constexpr int    NZ   = NzCount<1>::value;
                constexpr size_t smem = sizeof(double) * (size_t)(kTileCells + kTileGhostMax) * NZ;
                static bool      attr = false;
                if (!attr)
                  {
                    cudaFuncSetAttribute(k_assemble_gatherB2<NZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    attr = true;
                  }
                k_assemble_gatherB2<NZ><<<nt, kTileThreads, smem, s>>>(a, g);
              }
            ++*n_launches;
          }
        else if (n_rows > 0)
          {
            k_assemble_gatherB<<<grid_for(n_rows * 32, 256), 256, 0, s>>>(a, g);
            ++*n_launches;
          }
      }
    // boundary cells last (their rows hold the interior contributions by now)
    AsmArgs a   = a0;
    a.bnd_begin = 0;
    a.bnd_end   = a0.n_bnd;
    a.atomic_mode = 1;
    if (a.n_bnd > 0)
      {
        if (element == 0)
          k_assemble_boundary<0><<<grid_for(a.n_bnd, 64), 64, 0, s>>>(a, P);
        else
          k_assemble_boundary<1><<<grid_for(a.n_bnd, 64), 64, 0, s>>>(a, P);
        ++*n_launches;
      }
    return cudaGetLastError();
  }

  // one pass per colour (cells of a colour share no owned row); colours in ascending order
  cudaError_t
  launch_assemble(const int element, const AsmArgs &a0, const DevParams &P, const int n_colors,
                  const int64_t *color_begin, const int64_t *bnd_color_begin, cudaStream_t s, int *n_launches)
  {
    for (int col = 0; col < n_colors; ++col)
      {
        AsmArgs a   = a0;
        a.a_begin   = color_begin[col];
        a.a_end     = color_begin[col + 1];
        a.bnd_begin = bnd_color_begin[col];
        a.bnd_end   = bnd_color_begin[col + 1];
        if (element == 0)
          launch_color<0>(a, P, s, n_launches);
        else
          launch_color<1>(a, P, s, n_launches);
      }
    return cudaGetLastError();
  }

  cudaError_t
  launch_local_matrices(const int element, const AsmArgs &a, const DevParams &P, const int64_t first,
                        const int64_t n, double *out, double *rhs_out, cudaStream_t s)
  {
    if (n == 0)
      return cudaSuccess;
    const unsigned g = grid_for(n, 128);
    if (element == 0)
      {
        k_local_matrices<0, 1><<<g, 128, 0, s>>>(a, P, first, n, out, rhs_out);
        k_local_matrices<0, 2><<<g, 128, 0, s>>>(a, P, first, n, out, rhs_out);
        k_local_matrices<0, 4><<<g, 128, 0, s>>>(a, P, first, n, out, rhs_out);
      }
    else
      {
        k_local_matrices<1, 1><<<g, 128, 0, s>>>(a, P, first, n, out, rhs_out);
        k_local_matrices<1, 2><<<g, 128, 0, s>>>(a, P, first, n, out, rhs_out);
        k_local_matrices<1, 4><<<g, 128, 0, s>>>(a, P, first, n, out, rhs_out);
      }
    return cudaGetLastError();
  }

  cudaError_t
  launch_lognormal(const int64_t n, const int64_t *cell_ids, const int64_t first, const double sigma,
                   const uint64_t seed, double *k, cudaStream_t s)
  {
    if (n == 0)
      return cudaSuccess;
    k_lognormal<<<grid_for(n, 256), 256, 0, s>>>(n, cell_ids, first, sigma, seed, k);
    return cudaGetLastError();
  }

  // FP64 roofline denominator: 16 independent FMA chains per thread, no memory traffic
  __global__ void __launch_bounds__(256)
  k_fp64_peak(const int iters, const double a, const double b, double *__restrict__ sink)
  {
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)
      x[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int it = 0; it < iters; ++it)
      {
#pragma unroll
        for (int k = 0; k < 16; ++k)
          x[k] = fma(x[k], a, b);
      }
    double s = 0.;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      s += x[k];
    if (s == 12345.678) // never true: keeps the chains alive
      sink[0] = s;
  }
  cudaError_t
  launch_fp64_peak(double *sink, cudaEvent_t e0, cudaEvent_t e1, cudaStream_t s, double *seconds, double *flop)
  {
    const int    iters = 1 << 14, grid = 148 * 16, block = 256;
    k_fp64_peak<<<grid, block, 0, s>>>(256, 0.999999, 1e-7, sink); // warm-up
    cudaEventRecord(e0, s);
    k_fp64_peak<<<grid, block, 0, s>>>(iters, 0.999999, 1e-7, sink);
    cudaEventRecord(e1, s);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess)
      return e;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *seconds = ms * 1e-3;
    *flop    = 2. * 16. * (double)iters * (double)grid * (double)block;
    return cudaGetLastError();
  }

  cudaError_t
  launch_sum(const double *x, const size_t n, double *out, cudaStream_t s)
  {
    cudaMemsetAsync(out, 0, sizeof(double), s);
    k_sum<<<148 * 4, 256, 0, s>>>(x, n, out);
    return cudaGetLastError();
  }
} // namespace pe
