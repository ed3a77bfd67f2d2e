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
    std::vector<PeerPlan>   plan;
    std::vector<HaloPeer>  *peers = nullptr;
  };

  static CommState *
  cs(pe_handle *h)
  {
    return reinterpret_cast<CommState *>(h->comm);
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
    h->comm = reinterpret_cast<ncclComm *>(c);
    return PE_OK;
  }

  void
  comm_destroy(pe_handle *h)
  {
    if (!h->comm)
      return;
    CommState *c = cs(h);
    if (c->comm)
      ncclCommDestroy(c->comm);
    delete c;
    h->comm = nullptr;
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
    DevBuf<int64_t> d_cnt_mine, d_cnt_all;
    if (d_cnt_mine.alloc((size_t)W * 3) != cudaSuccess || d_cnt_all.alloc((size_t)W * W * 3) != cudaSuccess)
      return PE_ERR_ALLOC;
    cudaMemcpyAsync(d_cnt_mine.p, need_cnt.data(), sizeof(int64_t) * W * 3, cudaMemcpyHostToDevice, s);
    if (ncclAllGather(d_cnt_mine.p, d_cnt_all.p, (size_t)W * 3, ncclInt64, c->comm, s) != ncclSuccess)
      return PE_ERR_NCCL;
    std::vector<int64_t> cnt_all((size_t)W * W * 3);
    cudaMemcpyAsync(cnt_all.data(), d_cnt_all.p, sizeof(int64_t) * cnt_all.size(), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);

    // exchange the index lists (global ids)
    DevBuf<uint32_t> d_ghost;
    if (d_ghost.alloc(pt.ghost_cols.size()) != cudaSuccess)
      return PE_ERR_ALLOC;
    cudaMemcpyAsync(d_ghost.p, pt.ghost_cols.data(), sizeof(uint32_t) * pt.ghost_cols.size(), cudaMemcpyHostToDevice, s);
    std::vector<DevBuf<uint32_t>> d_want((size_t)W);
    std::vector<int64_t>          want_cnt((size_t)W, 0);
    for (int q = 0; q < W; ++q)
      {
        for (int b = 0; b < 3; ++b)
          want_cnt[q] += cnt_all[((size_t)q * W + me) * 3 + b];
        if (want_cnt[q] && d_want[q].alloc((size_t)want_cnt[q]) != cudaSuccess)
          return PE_ERR_ALLOC;
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
    if (ncclGroupEnd() != ncclSuccess)
      return PE_ERR_NCCL;
    cudaStreamSynchronize(s);

    h->peers.clear();
    c->plan.clear();
    for (int q = 0; q < W; ++q)
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
            cudaMemcpy(want.data(), d_want[q].p, sizeof(uint32_t) * want.size(), cudaMemcpyDeviceToHost);
            hp.send_rows.resize(want.size());
            for (size_t k = 0; k < want.size(); ++k)
              {
                const int64_t l = pt.global_to_owned(want[k]);
                if (l < 0)
                  {
                    h->err = "halo setup: a peer asked for a row this rank does not own";
                    return PE_ERR_NCCL;
                  }
                hp.send_rows[k] = (int32_t)l;
              }
            if (hp.d_send_rows.alloc(want.size()) != cudaSuccess || hp.d_send_buf.alloc(want.size()) != cudaSuccess ||
                hp.d_send_buf_xi.alloc((size_t)pl.send_cnt[1]) != cudaSuccess)
              return PE_ERR_ALLOC;
            cudaMemcpy(hp.d_send_rows.p, hp.send_rows.data(), sizeof(int32_t) * want.size(), cudaMemcpyHostToDevice);
          }
        c->plan.push_back(pl);
      }
    return PE_OK;
  }

  void
  comm_halo_exchange(void *ctx, double *x)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c || h->peers.empty())
      return;
    cudaStream_t s = h->stream;
    for (HaloPeer &hp : h->peers)
      if (!hp.send_rows.empty())
        {
          const int64_t n = (int64_t)hp.send_rows.size();
          k_gather<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, hp.d_send_rows.p, x, hp.d_send_buf.p);
          ++h->launches;
        }
    ncclGroupStart();
    for (size_t k = 0; k < h->peers.size(); ++k)
      {
        HaloPeer       &hp = h->peers[k];
        const PeerPlan &pl = c->plan[k];
        int64_t         so = 0;
        for (int b = 0; b < 3; ++b)
          {
            if (pl.send_cnt[b])
              ncclSend(hp.d_send_buf.p + so, (size_t)pl.send_cnt[b], ncclDouble, pl.rank, c->comm, s);
            so += pl.send_cnt[b];
            if (pl.recv_cnt[b])
              ncclRecv(x + h->L + pl.recv_off[b], (size_t)pl.recv_cnt[b], ncclDouble, pl.rank, c->comm, s);
          }
      }
    ncclGroupEnd();
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
    cudaStream_t    s = h->stream;
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
    ncclGroupStart();
    for (size_t q = 0; q < h->peers.size(); ++q)
      {
        HaloPeer       &hp = h->peers[q];
        const PeerPlan &pl = c->plan[q];
        int64_t         so = 0;
        for (int b = 0; b < 3; ++b)
          {
            if (pl.send_cnt[b])
              ncclSend(hp.d_send_buf.p + so, (size_t)pl.send_cnt[b], ncclDouble, pl.rank, c->comm, s);
            so += pl.send_cnt[b];
            if (pl.recv_cnt[b])
              ncclRecv(x + h->L + pl.recv_off[b], (size_t)pl.recv_cnt[b], ncclDouble, pl.rank, c->comm, s);
          }
        if (pl.send_cnt[1])
          ncclSend(hp.d_send_buf_xi.p, (size_t)pl.send_cnt[1], ncclDouble, pl.rank, c->comm, s);
        if (pl.recv_cnt[1])
          ncclRecv(x + k.L + k.H + k.Lx + (h->L + pl.recv_off[1] - k.hu_end), (size_t)pl.recv_cnt[1], ncclDouble, pl.rank,
                   c->comm, s);
      }
    ncclGroupEnd();
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
    cudaStream_t s  = h->stream;
    ncclGroupStart();
    for (int k = 0; k < n_comp; ++k)
      {
        double *q = p + (size_t)k * comp_stride;
        if (me > 0)
          {
            ncclSend(q + a * row_len, (size_t)row_len, ncclDouble, me - 1, c->comm, s);
            ncclRecv(q + (a - 1) * row_len, (size_t)row_len, ncclDouble, me - 1, c->comm, s);
          }
        if (me < W - 1)
          {
            ncclSend(q + (b - 1) * row_len, (size_t)row_len, ncclDouble, me + 1, c->comm, s);
            ncclRecv(q + b * row_len, (size_t)row_len, ncclDouble, me + 1, c->comm, s);
          }
      }
    ncclGroupEnd();
  }
  void
  comm_allgather_rows(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, const int64_t *bounds)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    const int    W = h->part.world;
    cudaStream_t s = h->stream;
    ncclGroupStart();
    for (int q = 0; q < W; ++q)
      for (int k = 0; k < n_comp; ++k)
        {
          double      *x = p + (size_t)k * comp_stride + bounds[q] * row_len;
          const size_t n = (size_t)((bounds[q + 1] - bounds[q]) * row_len);
          if (n)
            ncclBroadcast(x, x, n, ncclDouble, q, c->comm, s);
        }
    ncclGroupEnd();
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
    cudaStream_t s = h->stream;
    ncclGroupStart();
    for (int q = 0; q < W; ++q)
      for (int k = 0; k < n_comp; ++k)
        {
          const size_t off = (size_t)k * comp_stride + bounds[q] * row_len;
          const size_t n   = (size_t)((bounds[q + 1] - bounds[q]) * row_len);
          if (n)
            ncclReduce(in + off, out + off, n, ncclDouble, ncclSum, q, c->comm, s);
        }
    ncclGroupEnd();
  }

  void
  comm_allreduce2(void *ctx, const double *in, double *out, int n)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    ncclAllReduce(in, out, (size_t)n, ncclDouble, ncclSum, c->comm, h->stream);
  }

  void
  comm_allreduce(void *ctx, double *v, int n)
  {
    pe_handle *h = (pe_handle *)ctx;
    CommState *c = cs(h);
    if (!c)
      return;
    ncclAllReduce(v, v, (size_t)n, ncclDouble, ncclSum, c->comm, h->stream);
  }
} // namespace pe
