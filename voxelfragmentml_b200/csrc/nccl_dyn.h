// nccl_dyn.h — the handful of NCCL entry points the slab flood's exchange loop uses, resolved at run time from the libnccl.so.2 already in the
// process (a host application that brings its own NCCL, or the copy PyTorch loads) so that libvoxfrag.so carries no link-time dependency
// on a particular NCCL build.  Types come from <nccl.h>; nothing here is called unless a slab run asks for it.
#pragma once

#include <dlfcn.h>
#include <nccl.h>

struct VfNcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
};

inline const VfNcclApi* vf_nccl_api()
{
    static VfNcclApi api;
    static const VfNcclApi* ready = [] () -> const VfNcclApi* {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy that is loaded already, if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return nullptr;
        bool ok = true;
        auto sym = [&](const char* name) {
            void* p = dlsym(h, name);
            ok = ok && p != nullptr;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        return ok ? &api : nullptr;
    }();
    return ready;
}
