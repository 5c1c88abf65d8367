// libmlx5.so.1 look-alike: the direct-verbs entry points the backend resolves with dlsym().  The queue
// objects live in the mock libibverbs (mock_verbs.cc), exactly as libmlx5 is a provider plugged into
// libibverbs; this file only forwards.
#include <stdint.h>

#include "../verbs/abi/verbs_abi.h"

extern "C" int mock_mlx5dv_init_obj(rnabi::mlx5dv_obj* obj, uint64_t type);
extern "C" bool mock_mlx5dv_is_supported(rnabi::ibv_device* d);

extern "C" __attribute__((visibility("default"))) int mlx5dv_init_obj(rnabi::mlx5dv_obj* obj, uint64_t type) {
  return mock_mlx5dv_init_obj(obj, type);
}
extern "C" __attribute__((visibility("default"))) bool mlx5dv_is_supported(rnabi::ibv_device* d) {
  return mock_mlx5dv_is_supported(d);
}
