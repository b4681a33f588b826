// SFNO (FourCastNet-v2-small) step operator — placeholder until the SHT kernels land.
#include "engine.h"
namespace sky {
Engine* make_sfno_engine(const sky_sfno_config_t&, int) {
  set_error("SFNO engine not built yet");
  return nullptr;
}
}  // namespace sky
