// Device-resident replay buffer: gather of item slots into a collated batch (GPU collate).
//
// The rollout buffer (rift/gym_carla/buffer/cbv_rollout_buffer.py:16-138) keeps one slot per stored transition; every
// ragged feature tensor of a slot (agents, polygons, reference lines: first dimension varies per item) is stored zero
// padded to the arena's capacity for that dimension.  PlutoFeature.collate (pluto_feature.py:25-96) + the RL collates
// (rift_datamodule.py:20-51) zero-pad a mini-batch to ITS longest item: that is "copy the first n_batch rows of every
// selected slot", one contiguous byte range per (field, item).  One launch does it for all fields of the batch.
#include "common.cuh"
#include "../../include/rift_b200.h"

namespace rift {

struct GatherTable { rift_b200_gather_field f[RIFT_B200_GATHER_MAX_FIELDS]; };

__global__ void __launch_bounds__(256)
gather_fields_kernel(GatherTable t, const long long* __restrict__ idx, int bs) {
    pdl_grid_sync();
    const rift_b200_gather_field fd = t.f[blockIdx.y];
    for (int b = blockIdx.x; b < bs; b += gridDim.x) {
        const char* src = static_cast<const char*>(fd.src) + idx[b] * fd.item_stride_bytes;
        char* dst = static_cast<char*>(fd.dst) + (long long)b * fd.dst_stride_bytes;
        const long long n = fd.copy_bytes;
        if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)n) & 15) == 0) {
            const int4* s4 = reinterpret_cast<const int4*>(src);
            int4* d4 = reinterpret_cast<int4*>(dst);
            for (long long i = threadIdx.x; i < (n >> 4); i += 256) d4[i] = __ldg(s4 + i);
        } else {
            for (long long i = threadIdx.x; i < n; i += 256) dst[i] = src[i];
        }
    }
}

int launch_gather_fields(const rift_b200_gather_field* fields, int n_fields, const long long* idx, int bs, cudaStream_t st) {
    RIFT_REQUIRE(n_fields >= 0 && n_fields <= RIFT_B200_GATHER_MAX_FIELDS, "gather_fields: too many fields");
    if (n_fields == 0 || bs <= 0) return 0;
    GatherTable t;
    for (int i = 0; i < n_fields; ++i) {
        RIFT_REQUIRE(fields[i].src && fields[i].dst && fields[i].copy_bytes >= 0 && fields[i].copy_bytes <= fields[i].item_stride_bytes &&
                     fields[i].copy_bytes <= fields[i].dst_stride_bytes, "gather_fields: bad field descriptor");
        t.f[i] = fields[i];
    }
    launch_k(gather_fields_kernel, dim3(bs < 296 ? bs : 296, n_fields), 256, 0, st, t, idx, bs);
    RIFT_LAUNCH_OK();
    return 0;
}

}  // namespace rift

extern "C" int rift_b200_gather_fields(const rift_b200_gather_field* fields, int n_fields, const long long* idx, int bs,
                                       void* stream) {
    RIFT_REQUIRE(fields && idx, "gather_fields: null argument");
    return rift::launch_gather_fields(fields, n_fields, idx, bs, reinterpret_cast<cudaStream_t>(stream));
}
