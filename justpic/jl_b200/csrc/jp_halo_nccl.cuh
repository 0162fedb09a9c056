// jp_halo_nccl.cuh -- update_cell_halo! / update_halo! as library calls: pack kernels, the x -> y -> z schedule and the
// NCCL point-to-point transport (ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd) all live here, so that a host
// language only hands over a communicator (reference: src/CellArrays/ImplicitGlobalGrid.jl:36-41 = ImplicitGlobalGrid's
// update_halo! applied to each CellArray; scripts/temperature_advection3D_MPI.jl:83-91).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy already in the process when the host is PyTorch or
// NCCL.jl): the library has no link-time dependency on it and single-GPU users never load it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>          // types and enums only; every function is fetched with dlsym

struct JpNccl {
    void *h;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *);
    ncclResult_t (*CommCount)(const ncclComm_t, int *);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};

static JpNccl *jp_nccl() {
    static JpNccl api;
    static int state = 0;            // 0 untried, 1 ok, -1 failed
    if (state == 0) {
        state = -1;
        const char *names[] = {getenv("JP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (api.h) {
            bool ok = true;
#define JP_NCCL_SYM(field, sym) ok = ok && ((*(void **)(&api.field) = dlsym(api.h, sym)) != nullptr)
            JP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
            JP_NCCL_SYM(CommInitRank, "ncclCommInitRank");
            JP_NCCL_SYM(CommDestroy, "ncclCommDestroy");
            JP_NCCL_SYM(CommUserRank, "ncclCommUserRank");
            JP_NCCL_SYM(CommCount, "ncclCommCount");
            JP_NCCL_SYM(GroupStart, "ncclGroupStart");
            JP_NCCL_SYM(GroupEnd, "ncclGroupEnd");
            JP_NCCL_SYM(Send, "ncclSend");
            JP_NCCL_SYM(Recv, "ncclRecv");
            JP_NCCL_SYM(AllReduce, "ncclAllReduce");
            JP_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef JP_NCCL_SYM
            if (ok) state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}
#define JP_NCCL(call)                                                                               \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) return jp_fail(JP_ERR_CUDA, #call ": NCCL: %s", nc->GetErrorString(r__)); \
    } while (0)

extern "C" int jp_comm_unique_id(void *id128) {
    JpNccl *nc = jp_nccl();
    if (!nc) return jp_fail(JP_ERR_UNSUPPORTED, "jp_comm_unique_id: libnccl.so.2 could not be loaded (set JP_NCCL_LIB)");
    if (!id128) return jp_fail(JP_ERR_INVALID, "jp_comm_unique_id: null argument");
    ncclUniqueId id;
    JP_NCCL(nc->GetUniqueId(&id));
    memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return JP_OK;
}

extern "C" int jp_comm_init(const void *id128, int32_t nranks, int32_t rank, int32_t device, void **comm_out) {
    JpNccl *nc = jp_nccl();
    if (!nc) return jp_fail(JP_ERR_UNSUPPORTED, "jp_comm_init: libnccl.so.2 could not be loaded (set JP_NCCL_LIB)");
    if (!id128 || !comm_out || nranks < 1 || rank < 0 || rank >= nranks) return jp_fail(JP_ERR_INVALID, "jp_comm_init: bad argument");
    JP_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    JP_NCCL(nc->CommInitRank(&comm, nranks, id, rank));
    *comm_out = (void *)comm;
    return JP_OK;
}

extern "C" int jp_comm_destroy(void *comm) {
    JpNccl *nc = jp_nccl();
    if (!nc || !comm) return JP_OK;
    JP_NCCL(nc->CommDestroy((ncclComm_t)comm));
    return JP_OK;
}

// dt = MPI.Allreduce(maximum(abs.(V)), MPI.MAX) of the reference's MPI scripts (temperature_advection3D_MPI.jl:71): in place on n doubles
extern "C" int jp_allreduce_max(void *comm, double *buf, int32_t n, void *stream) {
    JpNccl *nc = jp_nccl();
    if (!nc) return jp_fail(JP_ERR_UNSUPPORTED, "jp_allreduce_max: libnccl.so.2 could not be loaded");
    if (!comm || !buf || n < 0) return jp_fail(JP_ERR_INVALID, "jp_allreduce_max: bad argument");
    JP_NCCL(nc->AllReduce(buf, buf, (size_t)n, ncclDouble, ncclMax, (ncclComm_t)comm, (cudaStream_t)stream));
    return JP_OK;
}

// ---- one cell-plane of every listed CellArray <-> one contiguous buffer.  grid.y = (array, slot) pair (the index bytes are
// array number arrs.n), grid.x * block = cells of the plane in memory order; y / z planes are contiguous runs in both source
// and buffer, an x plane is one element per row of nx cells (every element in its own 32-byte sector: the reason the default
// topology never splits x, halo.py).
template <bool PACK>
__global__ void __launch_bounds__(256) k_halo_plane(JpGrid g, int dim, int plane, HaloArrs arrs, uint8_t *index, unsigned char *buf, int M) {
    const int a = blockIdx.y / g.S, s = blockIdx.y % g.S;
    const int nx = g.n[0], ny = g.n[1];
    const int64_t slot_off = (int64_t)s * g.C;
    double *bd = (double *)buf + ((int64_t)a * g.S + s) * M;
    unsigned char *bb = buf + (int64_t)arrs.n * g.S * M * 8 + (int64_t)s * M;
    double *ad = a < arrs.n ? arrs.a[a] : nullptr;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        int64_t e;
        if (dim == 0) e = plane + (int64_t)nx * m;                                              // m = j + ny * k
        else if (dim == 1) { const int k = m / nx, i = m - k * nx; e = i + (int64_t)nx * (plane + (int64_t)ny * k); }
        else e = m + (int64_t)nx * ny * plane;
        e += slot_off;
        if (ad) { if (PACK) bd[m] = ad[e]; else ad[e] = bd[m]; }
        else    { if (PACK) bb[m] = index[e]; else index[e] = bb[m]; }
    }
}

// one plane of a plain grid array (x fastest, extents ext[]) <-> contiguous buffer (update_halo! of a velocity / vertex field)
template <bool PACK>
__global__ void __launch_bounds__(256) k_grid_plane(double *A, int e0, int e1, int e2, int dim, int plane, double *buf, int M) {
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        int64_t e;
        if (dim == 0) e = plane + (int64_t)e0 * m;
        else if (dim == 1) { const int k = m / e0, i = m - k * e0; e = i + (int64_t)e0 * (plane + (int64_t)e1 * k); }
        else e = m + (int64_t)e0 * e1 * plane;
        if (PACK) buf[m] = A[e]; else A[e] = buf[m];
    }
}

static int halo_buf_reserve(jp_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->halo_buf_bytes) return JP_OK;
    if (ctx->halo_buf) JP_CUDA(cudaFree(ctx->halo_buf));
    ctx->halo_buf = nullptr; ctx->halo_buf_bytes = 0;
    JP_CUDA(cudaMalloc(&ctx->halo_buf, bytes));
    ctx->halo_buf_bytes = bytes;
    return JP_OK;
}

// the exchange of one dimension: planes `send_lo` / `send_hi` go to the left / right neighbour, what arrives from them lands in
// planes `recv_lo` / `recv_hi`.  Two ranks that are each other's left AND right neighbour (periodic, two ranks along the
// dimension) post their sends as (to-left, to-right) and their receives as (from-right, from-left): NCCL matches several
// messages between one pair of ranks in posting order, so the plane sent to the left is the one the peer receives from its
// right.  A rank that is its own neighbour (periodic, one rank along the dimension) copies locally.
template <class PackFn, class UnpackFn>
static int halo_exchange_dim(JpNccl *nc, ncclComm_t comm, int me, int left, int right, size_t nb, unsigned char *buf, cudaStream_t st,
                             PackFn pack, UnpackFn unpack, int send_lo, int send_hi, int recv_lo, int recv_hi) {
    const size_t pitch = (nb + 255) & ~(size_t)255;
    unsigned char *sl = buf, *sr = buf + pitch, *rl = buf + 2 * pitch, *rr = buf + 3 * pitch;
    if (left >= 0) pack(send_lo, sl);
    if (right >= 0) pack(send_hi, sr);
    const bool self = me == -2 ? true : (left == me || right == me);     // no communicator: the only possible neighbour is the rank itself
    if (self) {
        if (left != right) return jp_fail(JP_ERR_INVALID, "jp_halo_exchange: a rank can only be its own neighbour on both sides (periodic, one rank along the dimension)");
        unpack(recv_hi, sl);
        unpack(recv_lo, sr);
        return JP_OK;
    }
    if (!nc || !comm) return jp_fail(JP_ERR_INVALID, "jp_halo_exchange: neighbours given but no communicator");
    JP_NCCL(nc->GroupStart());
    if (left >= 0) JP_NCCL(nc->Send(sl, nb, ncclUint8, left, comm, st));
    if (right >= 0) JP_NCCL(nc->Send(sr, nb, ncclUint8, right, comm, st));
    if (right >= 0) JP_NCCL(nc->Recv(rr, nb, ncclUint8, right, comm, st));
    if (left >= 0) JP_NCCL(nc->Recv(rl, nb, ncclUint8, left, comm, st));
    JP_NCCL(nc->GroupEnd());
    if (left >= 0) unpack(recv_lo, rl);
    if (right >= 0) unpack(recv_hi, rr);
    return JP_OK;
}

static int halo_comm_rank(JpNccl *nc, void *comm, int *me) {
    *me = -2;                                             // no communicator: every neighbour listed is the rank itself (periodic, undecomposed)
    if (comm) {
        if (!nc) return jp_fail(JP_ERR_UNSUPPORTED, "jp_halo_exchange: libnccl.so.2 could not be loaded");
        JP_NCCL(nc->CommUserRank((ncclComm_t)comm, me));
    }
    return JP_OK;
}

static void halo_mark_dirty(jp_ctx *ctx, int dim, int plane);

extern "C" int jp_halo_exchange(jp_ctx *ctx, void *comm, const int32_t *nbr, double *const *arrays, int32_t narrays, uint8_t *index,
                                void *stream) {
    if (!ctx || !nbr || !index) return jp_fail(JP_ERR_INVALID, "jp_halo_exchange: null argument");
    const JpGrid &g = ctx->g;
    if (narrays < 0 || narrays > JP_MAX_ARGS + 3) return jp_fail(JP_ERR_UNSUPPORTED, "jp_halo_exchange: too many arrays");
    HaloArrs h; h.n = narrays;
    for (int a = 0; a < narrays; a++) {
        if (!arrays || !arrays[a]) return jp_fail(JP_ERR_INVALID, "jp_halo_exchange: null array");
        h.a[a] = arrays[a];
    }
    JP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    JpNccl *nc = comm ? jp_nccl() : nullptr;
    int me;
    int rc = halo_comm_rank(nc, comm, &me);
    if (rc) return rc;
    for (int dim = 0; dim < g.ndim; dim++) {
        const int left = nbr[2 * dim], right = nbr[2 * dim + 1];
        if (left < 0 && right < 0) continue;
        const int n = g.n[dim];
        if (n < 4) return jp_fail(JP_ERR_INVALID, "jp_halo_exchange: a decomposed dimension needs at least 4 cells (overlap 2)");
        const int64_t M = plane_cells(g, dim);
        const size_t nb = (size_t)M * g.S * (8 * (size_t)narrays + 1);
        rc = halo_buf_reserve(ctx, 4 * ((nb + 255) & ~(size_t)255));
        if (rc) return rc;
        const dim3 grd((unsigned)((M + 255) / 256 < 64 ? (M + 255) / 256 : 64), (unsigned)((narrays + 1) * g.S));
        auto pack = [&](int plane, unsigned char *b) { k_halo_plane<true><<<grd, 256, 0, st>>>(g, dim, plane, h, index, b, (int)M); };
        auto unpack = [&](int plane, unsigned char *b) {
            mi_invalidate(ctx);
            k_halo_plane<false><<<grd, 256, 0, st>>>(g, dim, plane, h, index, b, (int)M);
            halo_mark_dirty(ctx, dim, plane);
        };
        // ImplicitGlobalGrid, overlap 2 / halo width 1 (1-based: plane 2 -> left's plane n, plane n-1 -> right's plane 1)
        rc = halo_exchange_dim(nc, (ncclComm_t)comm, me, left, right, nb, (unsigned char *)ctx->halo_buf, st, pack, unpack, 1, n - 2, 0, n - 1);
        if (rc) return rc;
        JP_CHECK_LAUNCH();
    }
    return JP_OK;
}

// update_halo!(A) of ImplicitGlobalGrid for a plain (staggered) grid array: extents ext[d] over the local base grid base[d]
// (= the cell grid the CellArrays live on); overlap ol = 2 + ext - base, so with halo width 1 (1-based) plane ol goes to the
// left neighbour's last plane and plane ext - ol + 1 to the right neighbour's first plane.  This is what carries the velocity
// ghost layers when V comes from a solver instead of an analytic formula.
extern "C" int jp_halo_exchange_grid(jp_ctx *ctx, void *comm, const int32_t *nbr, double *A, const int32_t *ext, void *stream) {
    if (!ctx || !nbr || !A || !ext) return jp_fail(JP_ERR_INVALID, "jp_halo_exchange_grid: null argument");
    const JpGrid &g = ctx->g;
    JP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    JpNccl *nc = comm ? jp_nccl() : nullptr;
    int me;
    int rc = halo_comm_rank(nc, comm, &me);
    if (rc) return rc;
    const int e0 = ext[0], e1 = ext[1], e2 = g.ndim == 3 ? ext[2] : 1;
    for (int dim = 0; dim < g.ndim; dim++) {
        const int left = nbr[2 * dim], right = nbr[2 * dim + 1];
        if (left < 0 && right < 0) continue;
        const int ol = 2 + ext[dim] - g.n[dim];
        if (ol < 2 || ext[dim] < 2 * ol) return jp_fail(JP_ERR_INVALID, "jp_halo_exchange_grid: array extent incompatible with overlap 2 on this grid");
        const int64_t M = (int64_t)e0 * e1 * e2 / ext[dim];
        const size_t nb = (size_t)M * 8;
        rc = halo_buf_reserve(ctx, 4 * ((nb + 255) & ~(size_t)255));
        if (rc) return rc;
        const unsigned blocks = (unsigned)((M + 255) / 256 < 1184 ? (M + 255) / 256 : 1184);
        auto pack = [&](int plane, unsigned char *b) { k_grid_plane<true><<<blocks, 256, 0, st>>>(A, e0, e1, e2, dim, plane, (double *)b, (int)M); };
        auto unpack = [&](int plane, unsigned char *b) { k_grid_plane<false><<<blocks, 256, 0, st>>>(A, e0, e1, e2, dim, plane, (double *)b, (int)M); };
        rc = halo_exchange_dim(nc, (ncclComm_t)comm, me, left, right, nb, (unsigned char *)ctx->halo_buf, st, pack, unpack,
                               ol - 1, ext[dim] - ol, 0, ext[dim] - 1);
        if (rc) return rc;
        JP_CHECK_LAUNCH();
    }
    return JP_OK;
}
