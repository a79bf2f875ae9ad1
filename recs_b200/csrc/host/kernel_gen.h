// Generic GPU query systems: source generation + NVRTC compilation (host only, no CUDA headers).
//
// The reference turns any Rust function whose arguments are `Read<T>` / `Write<T>` / `Entity` into a
// system (crates/ecs/src/systems/mod.rs:457-566, 627-739, 868-912) and iterates it over the
// concatenation of the matching archetypes' columns (crates/ecs/src/systems/iteration.rs:26-57,
// 151-245).  The GPU equivalent takes the body as CUDA C++ source, wraps it in a kernel that walks
// the same concatenation, and compiles it for sm_100a at registration time.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace recs {

struct KernelParamSpec {
    std::string type_name;  // C++ type of the element in the user's source
    uint32_t elem_size;     // sizeof(T), identical on host and device (POD components)
    uint32_t kind;          // RECS_KPARAM_READ / RECS_KPARAM_WRITE / RECS_KPARAM_ENTITY
};

struct KernelPlan {
    bool staged = false;       // pipelined form: column chunks of `chunk` entities cross HBM as TMA bulk copies
    uint32_t chunk = 256;      // entities per work item (pipelined) / per CTA pass (direct)
    uint32_t threads = 256;    // CTA size
    uint32_t smem_bytes = 0;   // dynamic shared memory of the pipelined form (all stages)
    std::vector<uint32_t> smem_offset;  // per parameter
    std::string source;        // user source + generated kernel `recs_generic_system`
};

// Decides the kernel form and generates the translation unit.
KernelPlan plan_kernel_system(const std::string& user_source, const std::string& body,
                              const std::vector<KernelParamSpec>& params, const std::string& uniform_type,
                              int force_direct);

// NVRTC (dlopen'ed: libnvrtc.so.12) -> cubin for sm_100a.  Needs no GPU.  Returns false and fills `log`.
bool compile_kernel_system(const std::string& source, std::vector<char>* cubin, std::string* log);

}  // namespace recs
