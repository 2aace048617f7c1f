#include "tad_nccl.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>

namespace tad {
namespace {

struct UniqueId { char internal[128]; };
typedef int (*fn_get_uid)(UniqueId *);
typedef int (*fn_init_rank)(void **, int, UniqueId, int);
typedef int (*fn_destroy)(void *);
typedef int (*fn_send)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_allgather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*fn_errstr)(int);

char g_err[256] = "";
void *g_lib = nullptr;

void *open_lib()
{
    if (g_lib) return g_lib;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_lib) return g_lib;
    }
    snprintf(g_err, sizeof(g_err), "cannot load libnccl.so.2: %s", dlerror());
    return nullptr;
}

void *sym(void *lib, const char *name)
{
    void *p = dlsym(lib, name);
    if (!p) snprintf(g_err, sizeof(g_err), "missing NCCL symbol %s", name);
    return p;
}

enum { F_INIT = 0, F_DESTROY, F_SEND, F_RECV, F_GSTART, F_GEND, F_ALLGATHER, F_ERRSTR };
constexpr int kNcclInt8 = 0;   // ncclInt8 / ncclChar

int check(NcclComm *c, int rc, const char *what)
{
    if (rc == 0) return 0;
    const char *s = c->fn[F_ERRSTR] ? ((fn_errstr)c->fn[F_ERRSTR])(rc) : "?";
    snprintf(g_err, sizeof(g_err), "%s failed: %s", what, s);
    return -1;
}

}  // namespace

const char *nccl_last_error() { return g_err; }

int nccl_get_unique_id(void *out, size_t bytes)
{
    if (bytes < sizeof(UniqueId)) return -1;
    void *lib = open_lib();
    if (!lib) return -1;
    fn_get_uid f = (fn_get_uid)sym(lib, "ncclGetUniqueId");
    if (!f) return -1;
    UniqueId id;
    if (f(&id) != 0) { snprintf(g_err, sizeof(g_err), "ncclGetUniqueId failed"); return -1; }
    memcpy(out, &id, sizeof(id));
    return 0;
}

int nccl_comm_init(NcclComm *c, int world, int rank, const void *unique_id, size_t bytes)
{
    if (!unique_id || bytes < sizeof(UniqueId)) { snprintf(g_err, sizeof(g_err), "missing ncclUniqueId"); return -1; }
    c->lib = open_lib();
    if (!c->lib) return -1;
    const char *names[8] = {"ncclCommInitRank", "ncclCommDestroy", "ncclSend", "ncclRecv", "ncclGroupStart",
                            "ncclGroupEnd", "ncclAllGather", "ncclGetErrorString"};
    for (int i = 0; i < 8; i++) {
        c->fn[i] = sym(c->lib, names[i]);
        if (!c->fn[i]) return -1;
    }
    UniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    c->world = world;
    c->rank = rank;
    return check(c, ((fn_init_rank)c->fn[F_INIT])(&c->comm, world, id, rank), "ncclCommInitRank");
}

void nccl_comm_destroy(NcclComm *c)
{
    if (c->comm && c->fn[F_DESTROY]) ((fn_destroy)c->fn[F_DESTROY])(c->comm);
    c->comm = nullptr;
}

int nccl_alltoallv(NcclComm *c, const void *send, const uint64_t *send_off, const uint64_t *send_bytes, void *recv,
                   const uint64_t *recv_off, const uint64_t *recv_bytes, cudaStream_t st)
{
    if (check(c, ((fn_group)c->fn[F_GSTART])(), "ncclGroupStart")) return -1;
    for (int p = 0; p < c->world; p++) {
        if (p == c->rank) continue;
        if (send_bytes[p] &&
            check(c, ((fn_send)c->fn[F_SEND])((const char *)send + send_off[p], send_bytes[p], kNcclInt8, p, c->comm, st), "ncclSend"))
            return -1;
        if (recv_bytes[p] &&
            check(c, ((fn_recv)c->fn[F_RECV])((char *)recv + recv_off[p], recv_bytes[p], kNcclInt8, p, c->comm, st), "ncclRecv"))
            return -1;
    }
    return check(c, ((fn_group)c->fn[F_GEND])(), "ncclGroupEnd");
}

int nccl_allgather(NcclComm *c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st)
{
    return check(c, ((fn_allgather)c->fn[F_ALLGATHER])(send, recv, bytes_per_rank, kNcclInt8, c->comm, st), "ncclAllGather");
}

}  // namespace tad
