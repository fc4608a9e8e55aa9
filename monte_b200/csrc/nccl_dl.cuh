// nccl_dl.cuh — the handful of NCCL entry points the multi-device paths use, bound with dlopen at first use
// (common.cu).  The declarations restate <nccl.h> (2.18+; checked against 2.27.3 here): opaque communicator,
// int result codes (0 = success), and the two enum values we pass.
#pragma once
#include "common.cuh"

typedef struct ncclComm *ncclComm_t;

namespace monte {

enum { NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT32 = 7 };   // ncclDataType_t
enum { NCCL_SUM = 0 };                                                          // ncclRedOp_t

struct NcclApi {
    bool loaded = false;
    int (*CommInitAll)(ncclComm_t *comms, int ndev, const int *devlist) = nullptr;
    int (*CommDestroy)(ncclComm_t comm) = nullptr;
    const char *(*GetErrorString)(int result) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void *send, void *recv, size_t count, int dtype, int op, int root, ncclComm_t comm, cudaStream_t st) = nullptr;
    int (*Send)(const void *send, size_t count, int dtype, int peer, ncclComm_t comm, cudaStream_t st) = nullptr;
    int (*Recv)(void *recv, size_t count, int dtype, int peer, ncclComm_t comm, cudaStream_t st) = nullptr;
    int (*GetVersion)(int *version) = nullptr;
};

const NcclApi *nccl_api();                  // nullptr (+ error text) if libnccl cannot be loaded
int nccl_comms(ncclComm_t **out);           // one communicator per bound device, rank = device index

#define MONTE_NCCL(call)                                                                        \
    do {                                                                                        \
        const int _r = (call);                                                                  \
        if (_r != 0) {                                                                          \
            ::monte::set_error("NCCL error %d (%s) in %s at %s:%d", _r,                         \
                               ::monte::nccl_api()->GetErrorString(_r), #call, __FILE__, __LINE__); \
            return MONTE_E_CUDA;                                                                \
        }                                                                                       \
    } while (0)

}  // namespace monte
