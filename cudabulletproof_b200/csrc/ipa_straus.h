// ipa_straus.h — host interface of ipa_straus.cu (the unfolded rounds of the inner-product argument).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cbp {
// largest number of base generators the composite rounds take (above, bpk_ipa_prove_device folds points and runs
// Pippenger MSMs until the vectors are this short): the crossover of the Straus window sums with Pippenger
constexpr size_t kIpaCompositeMax = 4096;
size_t ipa_composite_workspace_bytes(size_t mb);
int ipa_prove_composite(uint8_t* a, uint8_t* b, const uint8_t* g, const uint8_t* h, const uint8_t* Q, size_t mb,
                        int first_round, uint8_t* tr, uint8_t* u, uint8_t* ui, uint8_t* d_L, uint8_t* d_R,
                        uint8_t* d_x_out, uint8_t* ws, cudaStream_t st);
}  // namespace cbp
