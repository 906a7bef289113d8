// com.cu -- centres of mass of residues / segments on the device (SURVEY.md section
// 8(f) rank 2).
//
// Replaces the host einsum of the reference's center_of_mass
// (/root/reference/src/mdhelper/algorithm/molecule.py:15-310, used at
// analysis/structure.py:753-756 and :1485-1486) for entities that are consecutive runs
// of atoms: com_e = sum(m_a * r_a) / sum(m_a) in fp64, products rounded before they are
// added and atoms taken in index order -- the operation order of the host helper
// (numpy.bincount with weights), so the float32 results are bit-identical to it.

#include "common.cuh"

namespace {

// OUT = float: what capped_distance is given (it converts to float32); OUT = double: what
// the structure-factor classes keep in their float64 position buffer
template <typename OUT>
__global__ void com_kernel(const float *__restrict__ raw, int64_t frame_stride,
                           const int64_t *__restrict__ starts, const double *__restrict__ mass,
                           int64_t n_entities, OUT *__restrict__ out, int64_t out_stride)
{
    const int frame = blockIdx.y;
    const float *src = raw + (int64_t)frame * frame_stride;
    OUT *dst = out + (int64_t)frame * out_stride;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entities;
         e += (int64_t)gridDim.x * blockDim.x) {
        double sx = 0.0, sy = 0.0, sz = 0.0, sm = 0.0;
        for (int64_t a = starts[e]; a < starts[e + 1]; ++a) {
            const double m = mass[a];
            sx = __dadd_rn(sx, __dmul_rn(m, (double)src[3 * a]));
            sy = __dadd_rn(sy, __dmul_rn(m, (double)src[3 * a + 1]));
            sz = __dadd_rn(sz, __dmul_rn(m, (double)src[3 * a + 2]));
            sm = __dadd_rn(sm, m);
        }
        dst[3 * e] = (OUT)(sx / sm);
        dst[3 * e + 1] = (OUT)(sy / sm);
        dst[3 * e + 2] = (OUT)(sz / sm);
    }
}

}  // namespace

int com_configure_impl(mdh_ctx *c, int slot, int64_t n_atoms, int64_t n_entities,
                       const int64_t *starts, const double *masses)
{
    MDH_REQUIRE(slot >= 0 && slot < kComSlots, MDH_EINVAL, "com: slot must be in [0, %d)",
                kComSlots);
    MDH_REQUIRE(n_atoms >= 1 && n_entities >= 1 && n_entities <= n_atoms, MDH_EINVAL,
                "com: need 1 <= n_entities <= n_atoms");
    MDH_REQUIRE(starts && masses, MDH_EINVAL, "com: NULL argument");
    MDH_REQUIRE(starts[0] == 0 && starts[n_entities] == n_atoms, MDH_EINVAL,
                "com: starts must run from 0 to n_atoms");
    for (int64_t e = 0; e < n_entities; ++e)
        MDH_REQUIRE(starts[e] < starts[e + 1], MDH_EINVAL, "com: entity %lld is empty",
                    (long long)e);
    ComState &K = c->com[slot];
    K.configured = false;
    K.n_atoms = n_atoms; K.n_entities = n_entities;
    if (int rc = K.starts.reserve(sizeof(int64_t) * (n_entities + 1))) return rc;
    if (int rc = K.masses.reserve(sizeof(double) * n_atoms)) return rc;
    MDH_CUDA(cudaMemcpyAsync(K.starts.p, starts, sizeof(int64_t) * (n_entities + 1),
                             cudaMemcpyHostToDevice, c->stream));
    MDH_CUDA(cudaMemcpyAsync(K.masses.p, masses, sizeof(double) * n_atoms,
                             cudaMemcpyHostToDevice, c->stream));
    MDH_CUDA(cudaStreamSynchronize(c->stream));      // the sources are caller memory
    K.configured = true;
    return MDH_OK;
}

template <typename OUT>
static int com_reduce_any(mdh_ctx *c, int slot, const float *pos, int64_t stride, int location,
                          int n_frames, OUT *out_device, int64_t out_stride)
{
    MDH_REQUIRE(slot >= 0 && slot < kComSlots, MDH_EINVAL, "com: slot must be in [0, %d)",
                kComSlots);
    ComState &K = c->com[slot];
    MDH_REQUIRE(K.configured, MDH_ESTATE, "com: reduce before configure");
    MDH_REQUIRE(pos && out_device, MDH_EINVAL, "com: NULL argument");
    MDH_REQUIRE(n_frames >= 1 && n_frames <= 65535, MDH_EINVAL,
                "com: n_frames per call must be in [1, 65535]");
    MDH_REQUIRE(stride >= 3 * K.n_atoms, MDH_EINVAL, "com: frame_stride < 3*n_atoms");
    MDH_REQUIRE(out_stride >= 3 * K.n_entities, MDH_EINVAL,
                "com: out_frame_stride < 3*n_entities");
    MDH_REQUIRE(location == MDH_HOST || location == MDH_DEVICE, MDH_EINVAL,
                "com: invalid location");
    const float *dsrc = pos;
    int64_t dstride = stride;
    if (location == MDH_HOST) {
        if (int rc = K.raw.reserve(sizeof(float) * 3 * K.n_atoms * n_frames)) return rc;
        MDH_CUDA(mdh_copy_frames(K.raw.p, sizeof(float) * 3 * K.n_atoms, pos,
                                   sizeof(float) * stride, sizeof(float) * 3 * K.n_atoms,
                                   n_frames, cudaMemcpyHostToDevice, c->stream));
        dsrc = K.raw.as<float>();
        dstride = 3 * K.n_atoms;
    }
    dim3 grid((unsigned)std::min<int64_t>((K.n_entities + 127) / 128, 4096), n_frames);
    com_kernel<OUT><<<grid, 128, 0, c->stream>>>(dsrc, dstride, K.starts.as<int64_t>(),
                                                 K.masses.as<double>(), K.n_entities,
                                                 out_device, out_stride);
    MDH_CUDA(cudaGetLastError());
    c->launches++;
    return MDH_OK;
}

int com_reduce_impl(mdh_ctx *c, int slot, const float *pos, int64_t stride, int location,
                    int n_frames, float *out_device, int64_t out_stride)
{
    return com_reduce_any(c, slot, pos, stride, location, n_frames, out_device, out_stride);
}

int com_reduce_f64_impl(mdh_ctx *c, int slot, const float *pos, int64_t stride, int location,
                        int n_frames, double *out_device, int64_t out_stride)
{
    return com_reduce_any(c, slot, pos, stride, location, n_frames, out_device, out_stride);
}
