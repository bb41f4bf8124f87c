// comm.cu -- NCCL plumbing for row-sharded runs (one process per GPU, NVLink 5 / NVSwitch).
//
// Replaces the reference's only parallelism strategy, MultiDeviceKernel
// (/root/reference/gpytorch/kernels/multi_device_kernel.py:14-95: DataParallel scatter of x1 rows,
// CatLinearOperator gather of row blocks through one Python process) by: every rank owns a row block
// of K and of all CG vectors; per CG iteration one all-gather of the [n/g, 16] direction block and
// fp64 all-reduces of the packed dot-product messages.  libnccl is resolved at run time (dlopen of the
// libnccl.so.2 already loaded by torch, else the system one) so libgpbbmm has no link-time NCCL dependency.
#include <dlfcn.h>
#include <string.h>

#include "gp_common.cuh"

namespace gp {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_getuid)(ncclUniqueId_t*);
typedef int (*fn_initrank)(void**, int, ncclUniqueId_t, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);

static struct {
  void* h = nullptr;
  fn_getuid getuid = nullptr;
  fn_initrank initrank = nullptr;
  fn_destroy destroy = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_allgather allgather = nullptr;
  fn_errstr errstr = nullptr;
} g_nccl;

static int load_nccl() {
  if (g_nccl.h) return GP_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  GP_REQUIRE(g_nccl.h != nullptr, GP_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.getuid = (fn_getuid)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.initrank = (fn_initrank)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.allgather = (fn_allgather)dlsym(g_nccl.h, "ncclAllGather");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.h, "ncclGetErrorString");
  GP_REQUIRE(g_nccl.getuid && g_nccl.initrank && g_nccl.destroy && g_nccl.allreduce && g_nccl.allgather, GP_E_NCCL,
             "libnccl is missing required symbols");
  return GP_OK;
}

#define GP_NCCL(call)                                                                           \
  do {                                                                                          \
    int r__ = (call);                                                                           \
    if (r__ != 0) {                                                                             \
      set_error("%s:%d NCCL error %d: %s", __FILE__, __LINE__, r__, g_nccl.errstr ? g_nccl.errstr(r__) : "?"); \
      return GP_E_NCCL;                                                                         \
    }                                                                                           \
  } while (0)

// ncclDataType_t: ncclFloat32 = 7, ncclFloat64 = 8 ; ncclRedOp_t: ncclSum = 0
int nccl_allreduce_double(gp_comm* c, double* buf, size_t count, cudaStream_t st) {
  GP_NCCL(g_nccl.allreduce(buf, buf, count, 8, 0, c->nccl_comm, st));
  return GP_OK;
}
int nccl_allgather_float(gp_comm* c, float* buf, size_t count_per_rank, cudaStream_t st) {
  GP_NCCL(g_nccl.allgather(buf + (size_t)c->rank * count_per_rank, buf, count_per_rank, 7, c->nccl_comm, st));
  return GP_OK;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_comm_unique_id(uint8_t out[128]) {
  GP_CHECK(load_nccl());
  ncclUniqueId_t id;
  GP_NCCL(g_nccl.getuid(&id));
  memcpy(out, id.internal, 128);
  return GP_OK;
}

extern "C" int gp_comm_init(gp_comm** out, const uint8_t id[128], int rank, int world) {
  GP_REQUIRE(out && world >= 1 && rank >= 0 && rank < world, GP_E_SHAPE, "bad comm arguments");
  GP_CHECK(load_nccl());
  gp_comm* c = new gp_comm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId_t uid;
  memcpy(uid.internal, id, 128);
  int r = g_nccl.initrank(&c->nccl_comm, world, uid, rank);
  if (r != 0) {
    set_error("ncclCommInitRank failed: %d %s", r, g_nccl.errstr ? g_nccl.errstr(r) : "?");
    delete c;
    return GP_E_NCCL;
  }
  *out = c;
  return GP_OK;
}

extern "C" int gp_comm_destroy(gp_comm* c) {
  if (!c) return GP_OK;
  if (c->nccl_comm && g_nccl.destroy) g_nccl.destroy(c->nccl_comm);
  delete c;
  return GP_OK;
}
