// placeholder — replaced by the tcgen05 kernels
#include "s2s_internal.cuh"
namespace s2s {
void edge_embed_tc(const EdgeEmbedArgs&, cudaStream_t) { S2S_CHECK(false, "edge_embed_tc: not built"); }
void edge_transition_tc(const EdgeTransitionArgs&, cudaStream_t) { S2S_CHECK(false, "edge_transition_tc: not built"); }
}  // namespace s2s
