// Multi-GPU plumbing (one process per GPU): NCCL communicator owned by the handle, halo exchange
// of the interface DoFs for the distributed SpMV, all-reduce of the Krylov dot products.
// The reference is serial: this is new work defined by BASELINE north_star /
// SURVEY 8(e).  Ghost columns are kept sorted by global index, i.e. block-major [u | p_int | p_face]
// and by owner rank inside each block, so every (peer, block) pair is ONE contiguous slice of the
// halo part of a vector and is received in place.
#include "pe_internal.h"
#include "pe_k3.h"

#include <nccl.h>

#include <algorithm>
#include <cstring>

namespace pe
{
  namespace
  {
    __global__ void
    k_gather(const int64_t n, const int32_t *__restrict__ rows, const double *__restrict__ x, double *__restrict__ buf)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n)
        buf[i] = x[rows[i]];
    }

    // buf[i] = x[rows[i] + shift]
    __global__ void
    k_gather_shift(const int64_t n, const int32_t *__restrict__ rows, const int64_t shift, const double *__restrict__ x,
                   double *__restrict__ buf)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n)
        buf[i] = x[rows[i] + shift];
    }

    // rectangle [c0, c1) x [r0, r1) of a lexicographic grid with rows of row_len entries, n_comp components
    // comp_stride apart  <->  contiguous buffer [comp][row][col]
    __global__ void
    k_rect_pack(const double *__restrict__ grid, const int64_t row_len, const int64_t comp_stride, const int n_comp,
                const int c0, const int c1, const int r0, const int r1, double *__restrict__ buf)
    {
      const int64_t w = c1 - c0, n = w * (r1 - r0), t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n * n_comp)
        return;
      const int64_t k = t / n, q = t % n;
      buf[t]          = grid[k * comp_stride + (r0 + q / w) * row_len + c0 + q % w];
    }
    template <bool ADD>
    __global__ void
    k_rect_unpack(double *__restrict__ grid, const int64_t row_len, const int64_t comp_stride, const int n_comp, const int c0,
                  const int c1, const int r0, const int r1, const double *__restrict__ buf)
    {
      const int64_t w = c1 - c0, n = w * (r1 - r0), t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n * n_comp)
        return;
      const int64_t k = t / n, q = t % n;
      double       *d = grid + k * comp_stride + (r0 + q / w) * row_len + c0 + q % w;
      *d              = ADD ? *d + buf[t] : buf[t];
    }
    // zero rows [r0, r1) of every component
    __global__ void
    k_rows_zero(double *__restrict__ grid, const int64_t row_len, const int64_t comp_stride, const int n_comp, const int64_t r0,
                const int64_t r1)
    {
      const int64_t n = (r1 - r0) * row_len, t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n * n_comp)
        return;
      grid[(t / n) * comp_stride + r0 * row_len + t % n] = 0.;
    }

    struct Rect
    {
      int c0 = 0, c1 = 0, r0 = 0, r1 = 0;
      int64_t
      count() const
      {
        return (c1 > c0 && r1 > r0) ? (int64_t)(c1 - c0) * (r1 - r0) : 0;
      }
    };
    inline Rect
    clip_rows(Rect b, const int64_t ra, const int64_t rb)
    {
      b.r0 = (int)std::max<int64_t>(b.r0, ra);
      b.r1 = (int)std::min<int64_t>(b.r1, rb);
      return b;
    }
    // box <-> slab transfer plan of one grid (vertex or cell grid of one size)
    struct BoxPlan
    {
      int64_t            row_len = 0;
      int                n_comp  = 0;
      std::vector<Rect>  send, recv; // per rank (self included: local copy)
      std::vector<DevBuf<double>> sbuf, rbuf;
    };

    struct PeerPlan
    {
      int     rank;
      int64_t recv_off[3], recv_cnt[3]; // slices of the halo (relative to L)
      int64_t send_cnt[3];              // per block, in send_rows order
    };
  } // namespace

  struct CommState
  {
    ncclComm_t              comm = nullptr;
    ncclComm_t              comm2 = nullptr;          // duplicate communicator of lane 1 (pressure pipeline of the preconditioner)
    std::vector<PeerPlan>   plan;
    std::vector<HaloPeer>  *peers = nullptr;
    std::vector<int>        all_boxes;                // asm cell bounding box (i0, i1, j0, j1) of every rank
    std::vector<BoxPlan>    box_plans;                // [2 * kind + direction], built at first use
    ncclResult_t            first_err = ncclSuccess; // first failed NCCL call of a hot-path hook (latched)
    const char             *first_where = "";
    int                    *d_flag = nullptr;         // device int[2]: agreement of the ranks during setup
  };
  // The hot-path hooks are void (they run inside the captured iteration): a failed call is latched here and
  // turned into PE_ERR_NCCL by comm_status() at the synchronisation points of the solve.
#define PE_NCCL(c, call)                                             \
  do                                                                 \
    {                                                                \
      const ncclResult_t r_ = (call);                                \
      if (r_ != ncclSuccess && (c)->first_err == ncclSuccess)        \
        {                                                            \
          (c)->first_err   = r_;                                     \
          (c)->first_where = #call;                                  \
        }                                                            \
    }                                                                \
  while (0)

  static CommState *
  cs(pe_handle *h)
  {
    return reinterpret_cast<CommState *>(h->comm);
  }

  // Lanes: the displacement and the pressure pipeline of the preconditioner run concurrently on two streams; NCCL
  // operations of one communicator are serialised, so lane 1 has its own communicator (ncclCommSplit duplicate).
  // The host enqueues one lane at a time (h->cur_lane), in the same order on every rank.
  static inline cudaStream_t
  lane_stream(pe_handle *h)
  {
    return (h->cur_lane == 1 && h->stream2) ? h->stream2 : h->stream;
  }
  static inline ncclComm_t
  lane_comm(pe_handle *h)
  {
    CommState *c = cs(h);
    return (h->cur_lane == 1 && h->stream2 && c->comm2) ? c->comm2 : c->comm;
  }
  int
  comm_has_lanes(pe_handle *h)
  {
    CommState *c = cs(h);
    return c && c->comm2 && h->stream2;
  }

  int
  comm_unique_id(uint8_t id[128])
  {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    if (ncclGetUniqueId(&u) != ncclSuccess)
      return PE_ERR_NCCL;
    std::memcpy(id, &u, 128);
    return PE_OK;
  }

  int
  comm_init(pe_handle *h, const uint8_t id[128])
  {
    if (h->comm)
      return PE_OK;
    cudaSetDevice(h->device);
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    CommState *c = new CommState;
    if (ncclCommInitRank(&c->comm, h->cfg.world_size, u, h->cfg.rank) != ncclSuccess)
      {
        delete c;
        h->err = "ncclCommInitRank failed";
        return PE_ERR_NCCL;
      }
    // second lane: a duplicate communicator (all ranks, same order).  Failure is not fatal: one lane only.
    if (ncclCommSplit(c->comm, 0, h->cfg.rank, &c->comm2, nullptr) != ncclSuccess)
      c->comm2 = nullptr;
    if (cudaMalloc((void **)&c->d_flag, 2 * sizeof(int)) != cudaSuccess)
      {
        ncclCommDestroy(c->comm);
        delete c;
        h->err = "pe_comm_init: cudaMalloc failed";
        return PE_ERR_ALLOC;
      }
    h->comm = reinterpret_cast<ncclComm *>(c);
    return PE_OK;
  }

  void
  comm_destroy(pe_handle *h)
  {
    if (!h->comm)
      return;
    CommState *c = cs(h);
    if (c->d_flag)
      cudaFree(c->d_flag);
    if (c->comm2)
      ncclCommDestroy(c->comm2);
    if (c->comm)
      ncclCommDestroy(c->comm);
    delete c;
    h->comm = nullptr;
  }

  // PE_OK, or PE_ERR_NCCL with the text of the first failed call / the communicator's asynchronous error
  int
  comm_status(pe_handle *h)
  {
    CommState *c = cs(h);
    if (!c)
      return PE_OK;
    if (c->first_err != ncclSuccess)
      {
        h->err = std::string("NCCL call failed: ") + c->first_where + ": " + ncclGetErrorString(c->first_err);
        return PE_ERR_NCCL;
      }
    ncclResult_t async = ncclSuccess;
    if (ncclCommGetAsyncError(c->comm, &async) != ncclSuccess || async != ncclSuccess)
      {
        h->err = std::string("NCCL asynchronous error: ") + ncclGetErrorString(async);
        return PE_ERR_NCCL;
      }
    return PE_OK;
  }

  // true on every rank iff `mine` is true on every rank (collective; used where one rank must not leave a
  // collective sequence alone)
  static bool
  all_ranks_ok(pe_handle *h, const bool mine)
  {
    CommState *c = cs(h);
    const int  v = mine ? 1 : 0;
    int        r = 0;
    if (cudaMemcpyAsync(c->d_flag, &v, sizeof(int), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
      return false;
    if (ncclAllReduce(c->d_flag, c->d_flag + 1, 1, ncclInt, ncclMin, c->comm, h->stream) != ncclSuccess)
      return false;
    if (cudaMemcpyAsync(&r, c->d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess)
      return false;
    return r == 1;
  }

  // Tell every owner which of its rows we need; learn which of ours the others need.
  int
  comm_setup_halo(pe_handle *h)
  {
    if (!h->comm)
      {
        h->err = "world_size > 1 needs pe_comm_init before the hot path";
        return PE_ERR_NCCL;
      }
    CommState       *c  = cs(h);
    const Partition &pt = h->part;
    const int        W = pt.world, me = pt.rank;
    cudaStream_t     s = h->stream;
    // need[q][b] = slice of ghost_cols owned by q in block b
    std::vector<int64_t> need_off((size_t)W * 3, 0), need_cnt((size_t)W * 3, 0);
    for (int q = 0; q < W; ++q)
      for (int b = 0; b < 3; ++b)
        {
          const uint32_t lo = (uint32_t)pt.all_row_begin[(size_t)q * 3 + b], hi = (uint32_t)pt.all_row_end[(size_t)q * 3 + b];
          const auto     i0 = std::lower_bound(pt.ghost_cols.begin(), pt.ghost_cols.end(), lo);
          const auto     i1 = std::lower_bound(pt.ghost_cols.begin(), pt.ghost_cols.end(), hi);
          need_off[(size_t)q * 3 + b] = i0 - pt.ghost_cols.begin();
          need_cnt[(size_t)q * 3 + b] = q == me ? 0 : i1 - i0;
        }
    // all-gather the count matrix: cnt_all[r][q][b] = what r needs from q
    // COLLECTIVE: every rank of the partition must call this (pe_assemble_system / pe_set_solution after any
    // change of the mesh, pattern, boundary or permeability field).  No rank leaves the sequence alone: local
    // failures are agreed on with all_ranks_ok() before the next collective starts.
    DevBuf<int64_t> d_cnt_mine, d_cnt_all;
    DevBuf<uint32_t> d_ghost;
    bool ok = d_cnt_mine.alloc((size_t)W * 3) == cudaSuccess && d_cnt_all.alloc((size_t)W * W * 3) == cudaSuccess &&
              d_ghost.alloc(pt.ghost_cols.size()) == cudaSuccess;
    if (!all_ranks_ok(h, ok))
      {
        h->err = "halo setup: device allocation failed on a rank";
        return PE_ERR_ALLOC;
      }
    cudaMemcpyAsync(d_cnt_mine.p, need_cnt.data(), sizeof(int64_t) * W * 3, cudaMemcpyHostToDevice, s);
    ok = ncclAllGather(d_cnt_mine.p, d_cnt_all.p, (size_t)W * 3, ncclInt64, c->comm, s) == ncclSuccess;
    std::vector<int64_t> cnt_all((size_t)W * W * 3);
    ok = ok && cudaMemcpyAsync(cnt_all.data(), d_cnt_all.p, sizeof(int64_t) * cnt_all.size(), cudaMemcpyDeviceToHost, s) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(s) == cudaSuccess;

    // exchange the index lists (global ids)
    cudaMemcpyAsync(d_ghost.p, pt.ghost_cols.data(), sizeof(uint32_t) * pt.ghost_cols.size(), cudaMemcpyHostToDevice, s);
    std::vector<DevBuf<uint32_t>> d_want((size_t)W);
    std::vector<int64_t>          want_cnt((size_t)W, 0);
    if (ok)
      for (int q = 0; q < W; ++q)
        {
          for (int b = 0; b < 3; ++b)
            want_cnt[q] += cnt_all[((size_t)q * W + me) * 3 + b];
          if (want_cnt[q] && d_want[q].alloc((size_t)want_cnt[q]) != cudaSuccess)
            ok = false;
        }
    if (!all_ranks_ok(h, ok))
      {
        h->err = "halo setup: count exchange or allocation failed on a rank";
        return PE_ERR_NCCL;
      }
    ncclGroupStart();
    for (int q = 0; q < W; ++q)
      {
        if (q == me)
          continue;
        for (int b = 0; b < 3; ++b)
          if (need_cnt[(size_t)q * 3 + b])
            ncclSend(d_ghost.p + need_off[(size_t)q * 3 + b], (size_t)need_cnt[(size_t)q * 3 + b], ncclUint32, q, c->comm, s);
        int64_t off = 0;
        for (int b = 0; b < 3; ++b)
          {
            const int64_t n = cnt_all[((size_t)q * W + me) * 3 + b];
            if (n)
              ncclRecv(d_want[q].p + off, (size_t)n, ncclUint32, q, c->comm, s);
            off += n;
          }
      }
    ok = ncclGroupEnd() == ncclSuccess;
    ok = cudaStreamSynchronize(s) == cudaSuccess && ok;

    h->peers.clear();
    c->plan.clear();
    const char *why = "halo setup: index exchange failed on a rank";
    for (int q = 0; q < W && ok; ++q)
      {
        if (q == me)
          continue;
        int64_t nrecv = 0;
        for (int b = 0; b < 3; ++b)
          nrecv += need_cnt[(size_t)q * 3 + b];
        if (nrecv == 0 && want_cnt[q] == 0)
          continue;
        h->peers.emplace_back();
        HaloPeer &hp = h->peers.back();
        hp.rank      = q;
        PeerPlan pl{};
        pl.rank = q;
        for (int b = 0; b < 3; ++b)
          {
            pl.recv_off[b] = need_off[(size_t)q * 3 + b];
            pl.recv_cnt[b] = need_cnt[(size_t)q * 3 + b];
            pl.send_cnt[b] = cnt_all[((size_t)q * W + me) * 3 + b];
          }
        if (want_cnt[q])
          {
            std::vector<uint32_t> want((size_t)want_cnt[q]);
            if (cudaMemcpy(want.data(), d_want[q].p, sizeof(uint32_t) * want.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
              {
                ok = false;
                break;
              }
            hp.send_rows.resize(want.size());
            for (size_t k = 0; k < want.size(); ++k)
              {
                const int64_t l = pt.global_to_owned(want[k]);
                if (l < 0)
                  {
                    why = "halo setup: a peer asked for a row this rank does not own";
                    ok  = false;
                    break;
                  }
                hp.send_rows[k] = (int32_t)l;
              }
            if (!ok)
              break;
            if (hp.d_send_rows.alloc(want.size()) != cudaSuccess || hp.d_send_buf.alloc(want.size()) != cudaSuccess ||
                hp.d_send_buf_xi.alloc((size_t)pl.send_cnt[1]) != cudaSuccess ||
                cudaMemcpy(hp.d_send_rows.p, hp.send_rows.data(), sizeof(int32_t) * want.size(), cudaMemcpyHostToDevice) != cudaSuccess)
              {
                why = "halo setup: device allocation failed on a rank";
                ok  = false;
                break;
              }
          }
        c->plan.push_back(pl);
      }
    if (!all_ranks_ok(h, ok))
      {
        h->err = why;
        return PE_ERR_NCCL;
      }
    return PE_OK;
  }

  static void halo_exchange_blocks(pe_handle *h, double *x, int block_mask);
  void
  comm_halo_exchange(void *ctx, double *x)
  {
    halo_exchange_blocks((pe_handle *)ctx, x, 7);
  }
  // interior and face pressures only (lane 1 of the preconditioner must not read displacement entries that lane 0
  // is still writing)
  void
  comm_halo_exchange_p(void *ctx, double *x)
  {
    halo_exchange_blocks((pe_handle *)ctx, x, 6);
  }
  static void
  halo_exchange_blocks(pe_handle *h, double *x, const int block_mask)
  {
    CommState *c = cs(h);
    if (!c || h->peers.empty())
      return;
    cudaStream_t s = lane_stream(h);
    for (size_t k = 0; k < h->peers.size(); ++k)
      {
        HaloPeer       &hp = h->peers[k];
        const PeerPlan &pl = c->plan[k];
        if (hp.send_rows.empty())
          continue;
        // send rows are ordered by block [u | p_int | p_face]: gather the contiguous run of the wanted blocks
        int64_t o0 = 0, n = 0;
        for (int b = 0; b < 3; ++b)
          {
            if ((block_mask >> b) & 1)
              n += pl.send_cnt[b];
            else if (n == 0)
              o0 += pl.send_cnt[b];
          }
        if (n > 0)
          {
            k_gather<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, hp.d_send_rows.p + o0, x, hp.d_send_buf.p + o0);
            ++h->launches;
          }
      }
    PE_NCCL(c, ncclGroupStart());
    for (size_t k = 0; k < h->peers.size(); ++k)
      {
        HaloPeer       &hp = h->peers[k];
        const PeerPlan &pl = c->plan[k];
        int64_t         so = 0;
        for (int b = 0; b < 3; ++b)
          {
            const bool on = (block_mask >> b) & 1;
            if (on && pl.send_cnt[b])
              PE_NCCL(c, ncclSend(hp.d_send_buf.p + so, (size_t)pl.send_cnt[b], ncclDouble, pl.rank, lane_comm(h), s));
            so += pl.send_cnt[b];
            if (on && pl.recv_cnt[b])
              PE_NCCL(c, ncclRecv(x + h->L + pl.recv_off[b], (size_t)pl.recv_cnt[b], ncclDouble, pl.rank, lane_comm(h), s));
          }
      }
    PE_NCCL(c, ncclGroupEnd());
  }

  // three-field vectors (pe_k3.h): the (u, p) part as above plus the xi unknowns of the interior
  // pressures in the halo, in ONE NCCL group
  void
  comm_halo_exchange3(void *ctx, double *x)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c || h->peers.empty() || !h->k3)
      return;
    const K3System &k = *h->k3;
    cudaStream_t    s = lane_stream(h);
    for (size_t q = 0; q < h->peers.size(); ++q)
      {
        HaloPeer       &hp = h->peers[q];
        const PeerPlan &pl = c->plan[q];
        if (hp.send_rows.empty())
          continue;
        const int64_t n = (int64_t)hp.send_rows.size();
        k_gather<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, hp.d_send_rows.p, x, hp.d_send_buf.p);
        ++h->launches;
        const int64_t nx = pl.send_cnt[1];
        if (nx > 0)
          {
            // row r of the interior-pressure block -> xi at L + H + (r - pint_begin)
            k_gather_shift<<<(unsigned)((nx + 255) / 256), 256, 0, s>>>(nx, hp.d_send_rows.p + pl.send_cnt[0],
                                                                       k.L + k.H - k.pint_begin, x, hp.d_send_buf_xi.p);
            ++h->launches;
          }
      }
    PE_NCCL(c, ncclGroupStart());
    for (size_t q = 0; q < h->peers.size(); ++q)
      {
        HaloPeer       &hp = h->peers[q];
        const PeerPlan &pl = c->plan[q];
        int64_t         so = 0;
        for (int b = 0; b < 3; ++b)
          {
            if (pl.send_cnt[b])
              PE_NCCL(c, ncclSend(hp.d_send_buf.p + so, (size_t)pl.send_cnt[b], ncclDouble, pl.rank, lane_comm(h), s));
            so += pl.send_cnt[b];
            if (pl.recv_cnt[b])
              PE_NCCL(c, ncclRecv(x + h->L + pl.recv_off[b], (size_t)pl.recv_cnt[b], ncclDouble, pl.rank, lane_comm(h), s));
          }
        if (pl.send_cnt[1])
          PE_NCCL(c, ncclSend(hp.d_send_buf_xi.p, (size_t)pl.send_cnt[1], ncclDouble, pl.rank, lane_comm(h), s));
        if (pl.recv_cnt[1])
          PE_NCCL(c, ncclRecv(x + k.L + k.H + k.Lx + (h->L + pl.recv_off[1] - k.hu_end), (size_t)pl.recv_cnt[1], ncclDouble, pl.rank,
                   lane_comm(h), s));
      }
    PE_NCCL(c, ncclGroupEnd());
  }

  // only the xi unknowns of the halo (block-triangular preconditioner: B^T z_xi needs the neighbours' xi)
  void
  comm_halo_exchange_xi(void *ctx, double *x)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c || h->peers.empty() || !h->k3)
      return;
    const K3System &k = *h->k3;
    cudaStream_t    s = lane_stream(h);
    for (size_t q = 0; q < h->peers.size(); ++q)
      {
        HaloPeer       &hp = h->peers[q];
        const PeerPlan &pl = c->plan[q];
        const int64_t   nx = pl.send_cnt[1];
        if (hp.send_rows.empty() || nx <= 0)
          continue;
        k_gather_shift<<<(unsigned)((nx + 255) / 256), 256, 0, s>>>(nx, hp.d_send_rows.p + pl.send_cnt[0],
                                                                   k.L + k.H - k.pint_begin, x, hp.d_send_buf_xi.p);
        ++h->launches;
      }
    PE_NCCL(c, ncclGroupStart());
    for (size_t q = 0; q < h->peers.size(); ++q)
      {
        HaloPeer       &hp = h->peers[q];
        const PeerPlan &pl = c->plan[q];
        if (pl.send_cnt[1])
          PE_NCCL(c, ncclSend(hp.d_send_buf_xi.p, (size_t)pl.send_cnt[1], ncclDouble, pl.rank, lane_comm(h), s));
        if (pl.recv_cnt[1])
          PE_NCCL(c, ncclRecv(x + k.L + k.H + k.Lx + (h->L + pl.recv_off[1] - k.hu_end), (size_t)pl.recv_cnt[1], ncclDouble,
                              pl.rank, lane_comm(h), s));
      }
    PE_NCCL(c, ncclGroupEnd());
  }

  // ---- row slabs of the multigrid grids (GridComm hooks) ------------------------------------------------
  void
  comm_exchange_rows(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, int64_t a, int64_t b)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    const int    me = h->part.rank, W = h->part.world;
    cudaStream_t s  = lane_stream(h);
    PE_NCCL(c, ncclGroupStart());
    for (int k = 0; k < n_comp; ++k)
      {
        double *q = p + (size_t)k * comp_stride;
        if (me > 0)
          {
            PE_NCCL(c, ncclSend(q + a * row_len, (size_t)row_len, ncclDouble, me - 1, lane_comm(h), s));
            PE_NCCL(c, ncclRecv(q + (a - 1) * row_len, (size_t)row_len, ncclDouble, me - 1, lane_comm(h), s));
          }
        if (me < W - 1)
          {
            PE_NCCL(c, ncclSend(q + (b - 1) * row_len, (size_t)row_len, ncclDouble, me + 1, lane_comm(h), s));
            PE_NCCL(c, ncclRecv(q + b * row_len, (size_t)row_len, ncclDouble, me + 1, lane_comm(h), s));
          }
      }
    PE_NCCL(c, ncclGroupEnd());
  }
  void
  comm_allgather_rows(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, const int64_t *bounds)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    const int    W = h->part.world;
    cudaStream_t s = lane_stream(h);
    PE_NCCL(c, ncclGroupStart());
    for (int q = 0; q < W; ++q)
      for (int k = 0; k < n_comp; ++k)
        {
          double      *x = p + (size_t)k * comp_stride + bounds[q] * row_len;
          const size_t n = (size_t)((bounds[q + 1] - bounds[q]) * row_len);
          if (n)
            PE_NCCL(c, ncclBroadcast(x, x, n, ncclDouble, q, lane_comm(h), s));
        }
    PE_NCCL(c, ncclGroupEnd());
  }

  void
  comm_reduce_rows(void *ctx, const double *in, double *out, int64_t row_len, int n_comp, int64_t comp_stride,
                   const int64_t *bounds)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    const int    W = h->part.world;
    cudaStream_t s = lane_stream(h);
    PE_NCCL(c, ncclGroupStart());
    for (int q = 0; q < W; ++q)
      for (int k = 0; k < n_comp; ++k)
        {
          const size_t off = (size_t)k * comp_stride + bounds[q] * row_len;
          const size_t n   = (size_t)((bounds[q + 1] - bounds[q]) * row_len);
          if (n)
            PE_NCCL(c, ncclReduce(in + off, out + off, n, ncclDouble, ncclSum, q, lane_comm(h), s));
        }
    PE_NCCL(c, ncclGroupEnd());
  }

  // ---- box <-> slab transfers: the Biot vectors live on Morton blocks of cells (boxes), the fine multigrid levels
  // on row slabs.  Each rank sends / receives only the rectangles box x slab that overlap (round 1 reduced / all-
  // gathered the WHOLE grid on every rank: 268 MB per V-cycle on the 4096^2 vertex grid).
  // kind 0: vertex grid (row_len = N + 1; the box of cells [i0,i1) x [j0,j1) holds vertices [i0,i1] x [j0,j1]),
  // kind 1: cell grid (row_len = N).  `grow`: the box is widened by one entry on every side (the grid -> Biot
  // transfer of the cell grid reads the cells next to the box).
  int
  comm_set_box(pe_handle *h, const CellBox &box)
  {
    CommState *c = cs(h);
    if (!c)
      return PE_OK;
    const int W = h->part.world;
    DevBuf<int> d_mine, d_all;
    if (d_mine.alloc(4) != cudaSuccess || d_all.alloc((size_t)4 * W) != cudaSuccess)
      return PE_ERR_ALLOC;
    const int mine[4] = {box.i0, box.i1, box.j0, box.j1};
    cudaMemcpyAsync(d_mine.p, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream);
    if (ncclAllGather(d_mine.p, d_all.p, 4, ncclInt, lane_comm(h), lane_stream(h)) != ncclSuccess)
      return PE_ERR_NCCL;
    c->all_boxes.assign((size_t)4 * W, 0);
    cudaMemcpyAsync(c->all_boxes.data(), d_all.p, sizeof(int) * 4 * W, cudaMemcpyDeviceToHost, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess)
      return PE_ERR_CUDA;
    c->box_plans.clear();
    c->box_plans.resize(4);
    return PE_OK;
  }

  static BoxPlan *
  box_plan(pe_handle *h, const int kind, const int dir /* 0: boxes -> slabs, 1: slabs -> boxes */, const int64_t row_len,
           const int n_comp, const int64_t *bounds)
  {
    CommState *c = cs(h);
    if (c->all_boxes.empty() || c->box_plans.size() != 4)
      return nullptr;
    BoxPlan &P = c->box_plans[(size_t)(2 * kind + dir)];
    if (P.row_len == row_len && P.n_comp == n_comp && !P.send.empty())
      return &P;
    const int W = h->part.world, me = h->part.rank;
    const int n_rows = (int)bounds[W];
    auto box_of = [&](const int q) {
      const int *b = &c->all_boxes[(size_t)4 * q];
      Rect r;
      const int ext = kind == 0 ? 1 : 0; // vertices of the cell box
      const int grow = dir == 1 ? 1 : 0;
      r.c0 = std::max(0, b[0] - grow);
      r.c1 = (int)std::min<int64_t>(row_len, b[1] + ext + grow);
      r.r0 = std::max(0, b[2] - grow);
      r.r1 = std::min(n_rows, b[3] + ext + grow);
      if (b[1] <= b[0] || b[3] <= b[2])
        r = Rect{};
      return r;
    };
    P.row_len = row_len;
    P.n_comp  = n_comp;
    P.send.assign((size_t)W, Rect{});
    P.recv.assign((size_t)W, Rect{});
    P.sbuf.clear();
    P.rbuf.clear();
    P.sbuf.resize((size_t)W);
    P.rbuf.resize((size_t)W);
    for (int q = 0; q < W; ++q)
      {
        if (dir == 0)
          {
            P.send[(size_t)q] = clip_rows(box_of(me), bounds[q], bounds[q + 1]); // my box x slab of q
            P.recv[(size_t)q] = clip_rows(box_of(q), bounds[me], bounds[me + 1]); // box of q x my slab
          }
        else
          {
            P.send[(size_t)q] = clip_rows(box_of(q), bounds[me], bounds[me + 1]);
            P.recv[(size_t)q] = clip_rows(box_of(me), bounds[q], bounds[q + 1]);
          }
        if (q != me)
          {
            if (P.sbuf[(size_t)q].alloc((size_t)(P.send[(size_t)q].count() * n_comp)) != cudaSuccess ||
                P.rbuf[(size_t)q].alloc((size_t)(P.recv[(size_t)q].count() * n_comp)) != cudaSuccess)
              return nullptr;
          }
      }
    return &P;
  }

  static void
  box_transfer(pe_handle *h, BoxPlan &P, const double *in, double *out, const int64_t comp_stride, const bool add)
  {
    CommState   *c  = cs(h);
    const int    W  = h->part.world, me = h->part.rank;
    cudaStream_t s  = lane_stream(h);
    auto grid_of = [](const int64_t n) { return (unsigned)((n + 255) / 256); };
    for (int q = 0; q < W; ++q)
      {
        const Rect   &r = P.send[(size_t)q];
        const int64_t n = r.count() * P.n_comp;
        if (q == me || n == 0)
          continue;
        k_rect_pack<<<grid_of(n), 256, 0, s>>>(in, P.row_len, comp_stride, P.n_comp, r.c0, r.c1, r.r0, r.r1, P.sbuf[(size_t)q].p);
        ++h->launches;
      }
    PE_NCCL(c, ncclGroupStart());
    for (int q = 0; q < W; ++q)
      {
        if (q == me)
          continue;
        const int64_t ns = P.send[(size_t)q].count() * P.n_comp, nr = P.recv[(size_t)q].count() * P.n_comp;
        if (ns)
          PE_NCCL(c, ncclSend(P.sbuf[(size_t)q].p, (size_t)ns, ncclDouble, q, lane_comm(h), s));
        if (nr)
          PE_NCCL(c, ncclRecv(P.rbuf[(size_t)q].p, (size_t)nr, ncclDouble, q, lane_comm(h), s));
      }
    PE_NCCL(c, ncclGroupEnd());
    // own part: in -> out directly (in and out may be the same array: then nothing to do)
    {
      const Rect   &r = P.send[(size_t)me];
      const int64_t n = r.count() * P.n_comp;
      if (n && in != out)
        {
          // pack + unpack through a temporary is avoided: one kernel reads `in` and writes / adds into `out`
          if (P.sbuf[(size_t)me].n < (size_t)n)
            P.sbuf[(size_t)me].alloc((size_t)n);
          k_rect_pack<<<grid_of(n), 256, 0, s>>>(in, P.row_len, comp_stride, P.n_comp, r.c0, r.c1, r.r0, r.r1, P.sbuf[(size_t)me].p);
          if (add)
            k_rect_unpack<true><<<grid_of(n), 256, 0, s>>>(out, P.row_len, comp_stride, P.n_comp, r.c0, r.c1, r.r0, r.r1, P.sbuf[(size_t)me].p);
          else
            k_rect_unpack<false><<<grid_of(n), 256, 0, s>>>(out, P.row_len, comp_stride, P.n_comp, r.c0, r.c1, r.r0, r.r1, P.sbuf[(size_t)me].p);
          h->launches += 2;
        }
    }
    for (int q = 0; q < W; ++q)
      {
        const Rect   &r = P.recv[(size_t)q];
        const int64_t n = r.count() * P.n_comp;
        if (q == me || n == 0)
          continue;
        if (add)
          k_rect_unpack<true><<<grid_of(n), 256, 0, s>>>(out, P.row_len, comp_stride, P.n_comp, r.c0, r.c1, r.r0, r.r1, P.rbuf[(size_t)q].p);
        else
          k_rect_unpack<false><<<grid_of(n), 256, 0, s>>>(out, P.row_len, comp_stride, P.n_comp, r.c0, r.c1, r.r0, r.r1, P.rbuf[(size_t)q].p);
        ++h->launches;
      }
  }

  // out[my slab] = sum over the ranks of in[their box]  (in: this rank's contribution, valid on its box)
  void
  comm_boxes_to_slabs(void *ctx, const double *in, double *out, int64_t row_len, int n_comp, int64_t comp_stride,
                      const int64_t *bounds, int kind)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    BoxPlan *P = box_plan(h, kind, 0, row_len, n_comp, bounds);
    if (!P)
      {
        comm_reduce_rows(ctx, in, out, row_len, n_comp, comp_stride, bounds); // no box information: whole slabs
        return;
      }
    const int     me = h->part.rank;
    const int64_t n  = (bounds[me + 1] - bounds[me]) * row_len * n_comp;
    if (n > 0)
      {
        k_rows_zero<<<(unsigned)((n + 255) / 256), 256, 0, lane_stream(h)>>>(out, row_len, comp_stride, n_comp, bounds[me], bounds[me + 1]);
        ++h->launches;
      }
    box_transfer(h, *P, in, out, comp_stride, true);
  }
  // p[my box, widened by one] = the slab owners' values  (p: my slab rows valid on entry)
  void
  comm_slabs_to_boxes(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, const int64_t *bounds, int kind)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    BoxPlan *P = box_plan(h, kind, 1, row_len, n_comp, bounds);
    if (!P)
      {
        comm_allgather_rows(ctx, p, row_len, n_comp, comp_stride, bounds);
        return;
      }
    box_transfer(h, *P, p, p, comp_stride, false);
  }

  void
  comm_allreduce2(void *ctx, const double *in, double *out, int n)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    PE_NCCL(c, ncclAllReduce(in, out, (size_t)n, ncclDouble, ncclSum, lane_comm(h), lane_stream(h)));
  }

  void
  comm_allreduce_max(void *ctx, double *v, int n)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    PE_NCCL(c, ncclAllReduce(v, v, (size_t)n, ncclDouble, ncclMax, lane_comm(h), lane_stream(h)));
  }

  void
  comm_allreduce(void *ctx, double *v, int n)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    PE_NCCL(c, ncclAllReduce(v, v, (size_t)n, ncclDouble, ncclSum, lane_comm(h), lane_stream(h)));
  }
} // namespace pe
