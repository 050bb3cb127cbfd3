// tcgen05 path of the DiffWave layer (placeholder until the kernels land).
#include "../../include/sddm_b200.h"
#include "diffwave.cuh"

namespace sddm {
int launch_dw_layer_tc(const DwLayerTc&, cudaStream_t) { set_error("DiffWave tcgen05 path not built"); return SDDM_E_INVALID; }
int launch_dw_cond_tc(const DwCondTc&, cudaStream_t) { set_error("DiffWave tcgen05 path not built"); return SDDM_E_INVALID; }
}  // namespace sddm
