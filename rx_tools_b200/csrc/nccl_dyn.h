// nccl_dyn.h — NCCL bound at run time (dlopen "libnccl.so.2"), so that librxb200.so has no link-time
// dependency on it: single-GPU users never load NCCL, a process that already has one (torch) shares it.
// Only the handful of entry points the rx_power collation needs (ONE all-gather of the
// spectrum rows per report, src/rtl_power.c:1047-1050 being the consumer).
#pragma once
#include <stddef.h>
#include <cuda_runtime.h>
#if __has_include(<nccl.h>)
#include <nccl.h>
#else
// minimal declarations matching NCCL 2.x's public ABI
typedef struct ncclComm *ncclComm_t;
#define NCCL_UNIQUE_ID_BYTES 128
typedef struct { char internal[NCCL_UNIQUE_ID_BYTES]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 } ncclDataType_t;
#endif

namespace rxb {

struct NcclApi {
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)(void);
	ncclResult_t (*GroupEnd)(void);
	const char *(*GetErrorString)(ncclResult_t);
	ncclResult_t (*GetVersion)(int *);
};

// nullptr (with the thread's error text set) when libnccl.so.2 cannot be loaded
const NcclApi *nccl_api();

}  // namespace rxb
