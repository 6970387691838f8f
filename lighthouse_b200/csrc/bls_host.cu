// bls_host.cu — placeholder until the BLS path lands.
#include "ctx.h"
namespace lhb200 {
int32_t bls_init() { return LHB200_OK; }
void bls_shutdown() {}
}  // namespace lhb200
