// ABI bookkeeping for include/dgcnn_b200.h.
#include "common.cuh"

extern "C" int dgcnn_abi_version(void) { return DGCNN_B200_ABI_VERSION; }

extern "C" const char* dgcnn_status_string(int status) {
    switch (status) {
        case DGCNN_OK: return "ok";
        case DGCNN_ERR_INVALID_ARGUMENT: return "invalid argument";
        case DGCNN_ERR_UNSUPPORTED: return "size not supported by the kernels";
        case DGCNN_ERR_WORKSPACE: return "workspace missing or too small";
        case DGCNN_ERR_CUDA: return "CUDA launch failed";
        default: return "unknown status";
    }
}
