// pbd_tile.cu -- tile backend (placeholder until the persistent kernel lands).
#include "pbd_body.h"
namespace pbd {
Backend* make_tile_backend(const pbd_options&, int) { return nullptr; }
}  // namespace pbd
