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

struct recs_gpu_kernel_param;

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

// A `Commands` parameter: type_name / elem_size describe the create() record (elem_size 0: remove() only).
KernelParamSpec commands_param_spec(const recs_gpu_kernel_param& p);

// Decides the kernel form and generates the translation unit.
KernelPlan plan_kernel_system(const std::string& user_source, const std::string& body,
                              const std::vector<KernelParamSpec>& params, const std::string& uniform_type,
                              int force_direct);

// Systems with a nested `Query<(...)>` parameter (crates/ecs/src/systems/mod.rs:915-934, 1028-1140; the reference's
// `gravity` and `collision`): for every entity of the outer parameters the nested query is iterated IN QUERY ORDER and
// folded into an accumulator --
//     ACC acc = ACC();  for each query entity: pair(acc, outer reads..., query params...);  finish(outer params..., acc);
// Query entities are staged through shared memory in tiles of 256; one thread per outer entity, so a body that adds
// in the order it is called reproduces the reference's sequential `.sum()` bit for bit.
KernelPlan plan_pair_kernel_system(const std::string& user_source, const std::string& acc_type, const std::string& pair_fn,
                                   const std::string& finish_fn, const std::vector<KernelParamSpec>& outer,
                                   const std::vector<KernelParamSpec>& query, const std::string& uniform_type);

// NVRTC (dlopen'ed: libnvrtc.so.12) -> cubin for sm_100a.  Needs no GPU.  Returns false and fills `log`.
bool compile_kernel_system(const std::string& source, std::vector<char>* cubin, std::string* log);

}  // namespace recs
