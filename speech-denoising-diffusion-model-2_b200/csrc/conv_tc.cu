// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
#include "../../include/sddm_b200.h"
namespace sddm {
bool conv_tc_supported(const ConvP&) { return false; }
int conv_tc_nparts(int Hout, int Wout) { return (Hout / 16) * (Wout / 8) * 4; }
int launch_conv_tc(const ConvP&, cudaStream_t) { set_error("tcgen05 conv not built"); return SDDM_E_INVALID; }
}
extern "C" int sddm_debug_umma_probe(int, int, int, float*) { sddm::set_error("tcgen05 probe not built"); return SDDM_E_INVALID; }
