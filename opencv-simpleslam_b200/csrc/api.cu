// api.cu - error reporting, weight-blob parsing and misc C-ABI entry points of libb200slam.
#include <cstdlib>
#include "common.cuh"

#include <cstdarg>

namespace b2s {

static thread_local char g_err[1024] = "";

bool pdl_enabled() {
  static const bool on = [] { const char* e = std::getenv("B2S_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct BlobEntry {   // 128 bytes, see weights.py::pack_state
  char name[96];
  uint32_t ndim;
  uint32_t dims[4];
  uint32_t reserved;
  uint64_t offset;
};
static_assert(sizeof(BlobEntry) == 128, "blob entry layout");

int WeightBlob::parse(const void* blob, size_t nbytes) {
  const uint8_t* b = static_cast<const uint8_t*>(blob);
  if (nbytes < 16 || std::memcmp(b, "B2SW", 4) != 0) { set_error("weight blob: bad magic"); return B2S_EINVAL; }
  uint32_t version, n;
  std::memcpy(&version, b + 4, 4);
  std::memcpy(&n, b + 8, 4);
  if (version != 1 || 16 + (size_t)n * sizeof(BlobEntry) > nbytes) { set_error("weight blob: bad header (version %u, %u tensors)", version, n); return B2S_EINVAL; }
  for (uint32_t i = 0; i < n; ++i) {
    BlobEntry e;
    std::memcpy(&e, b + 16 + (size_t)i * sizeof(BlobEntry), sizeof(BlobEntry));
    e.name[95] = 0;
    TensorView v;
    v.ndim = (int)e.ndim;
    size_t numel = 1;
    for (int d = 0; d < 4; ++d) { v.dims[d] = (int)e.dims[d]; numel *= e.dims[d]; }
    if (e.ndim > 4 || (e.offset & 3) || e.offset + numel * sizeof(float) > nbytes) { set_error("weight blob: tensor %s out of bounds", e.name); return B2S_EINVAL; }
    v.data = reinterpret_cast<const float*>(b + e.offset);
    t[std::string(e.name)] = v;
  }
  return 0;
}

const TensorView* WeightBlob::get(const std::string& name, size_t expect) const {
  auto it = t.find(name);
  if (it == t.end()) { set_error("weight blob: missing tensor '%s'", name.c_str()); return nullptr; }
  if (expect && it->second.numel() != expect) {
    set_error("weight blob: tensor '%s' has %zu elements, expected %zu", name.c_str(), it->second.numel(), expect);
    return nullptr;
  }
  return &it->second;
}

}  // namespace b2s

extern "C" int b2s_version(void) { return B2S_VERSION; }
extern "C" const char* b2s_last_error(void) { return b2s::g_err; }
extern "C" int b2s_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
