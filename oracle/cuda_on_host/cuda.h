// forwards to the host meaning of the CUDA runtime (oracle/cuda_on_host/cuda_on_host.h); test infrastructure
#include "cuda_on_host.h"
