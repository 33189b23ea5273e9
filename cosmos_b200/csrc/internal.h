// Declarations shared by the .cu files behind the C ABI (include/cosmos_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cosmos_b200.h"

namespace cb {

typedef cosmos_ema_chunk EmaChunk;

cudaError_t launch_ema(const EmaChunk* table, int n_chunks, double momentum, int dtype, int sm_count, cudaStream_t stream);

}  // namespace cb
