// pbd_tileplan.cpp -- tile schedule builder (placeholder until the tile backend lands).
#include "pbd_plan.h"
namespace pbd {
bool build_tile_plan(const MeshView&, const pbd_options&, uint32_t, uint32_t, Plan&, std::string& err) {
  err = "tile backend not built yet";
  return false;
}
}  // namespace pbd
