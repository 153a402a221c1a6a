// comm.cu -- the one collective of the path: NCCL reduce-scatter of partial output spectra for the input-sharded
// MIMO engine ("C5 MIMO, input-sharded").  libnccl is loaded at run time (dlopen), so libbbx has no
// link-time dependency on it and every other entry point works without NCCL installed.  The host application
// creates the communicator: rank 0 calls bbx_comm_unique_id(), ships the 128 bytes to the other ranks by any means
// (bench.py / the tests use torch.distributed), every rank calls bbx_comm_create().
#include <dlfcn.h>

#include "common.cuh"

namespace bbx {

struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy = (int (*)(NcclComm))dlsym(api.lib, "ncclCommDestroy");
      api.ReduceScatter =
          (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.lib, "ncclReduceScatter");
      api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
      api.GetVersion = (int (*)(int*))dlsym(api.lib, "ncclGetVersion");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.ReduceScatter) {
        dlclose(api.lib);
        api.lib = nullptr;
      }
    }
  }
  return api.lib ? &api : nullptr;
}

#define BBX_NCCL_TRY(api, expr)                                                                       \
  do {                                                                                                \
    int _r = (expr);                                                                                  \
    if (_r != 0) {                                                                                    \
      bbx::set_error("%s failed: %s", #expr, (api)->GetErrorString ? (api)->GetErrorString(_r) : "?"); \
      return BBX_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)

}  // namespace bbx

using namespace bbx;

struct bbx_comm {
  NcclComm nccl = nullptr;
  int world = 1, rank = 0, device = 0;
};

namespace bbx {
// sum over ranks of `send` ([world][recvcount] floats), rank r receives chunk r
int comm_reduce_scatter_f32(bbx_comm* c, const float* send, float* recv, size_t recvcount, cudaStream_t st) {
  NcclApi* api = nccl_api();
  BBX_REQUIRE(api && c && c->nccl, "reduce-scatter without a communicator");
  BBX_NCCL_TRY(api, api->ReduceScatter(send, recv, recvcount, /*ncclFloat32*/ 7, /*ncclSum*/ 0, c->nccl, st));
  return BBX_OK;
}
int comm_world(const bbx_comm* c) { return c ? c->world : 1; }
int comm_rank(const bbx_comm* c) { return c ? c->rank : 0; }
}  // namespace bbx

extern "C" {

int bbx_comm_available(void) { return nccl_api() ? 1 : 0; }

int bbx_comm_unique_id(uint8_t* id128) {
  BBX_REQUIRE(id128 != nullptr, "bbx_comm_unique_id: null buffer");
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded");
    return BBX_ERR_UNSUPPORTED;
  }
  NcclUniqueId id;
  BBX_NCCL_TRY(api, api->GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return BBX_OK;
}

int bbx_comm_create(int world, int rank, const uint8_t* id128, int device, bbx_comm** out) {
  BBX_REQUIRE(out && id128 && world >= 1 && rank >= 0 && rank < world, "bbx_comm_create: bad argument");
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded");
    return BBX_ERR_UNSUPPORTED;
  }
  int rc = require_device();
  if (rc) return rc;
  DeviceGuard dg(device);
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  bbx_comm* c = new bbx_comm();
  c->world = world;
  c->rank = rank;
  c->device = device;
  int r = api->CommInitRank(&c->nccl, world, id, rank);
  if (r != 0) {
    set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    delete c;
    return BBX_ERR_CUDA;
  }
  *out = c;
  return BBX_OK;
}

int bbx_comm_destroy(bbx_comm* c) {
  if (!c) return BBX_OK;
  NcclApi* api = nccl_api();
  if (api && c->nccl) {
    DeviceGuard dg(c->device);
    api->CommDestroy(c->nccl);
  }
  delete c;
  return BBX_OK;
}

}  // extern "C"
