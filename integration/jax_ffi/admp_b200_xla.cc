// XLA typed-FFI shim over the C ABI of libadmp_b200.so (include/admp_b200.h): one handler per compute entry point, so that the
// kernels run as jax.ffi custom calls on JAX's own CUDA stream and device buffers.
//
// STATUS: UNTESTED. JAX / jaxlib (and therefore xla/ffi/api/ffi.h) are not installable in the build image of this repository
// (no wheels, no network); the file is the binding a maintainer adds on a machine that has them. The tested binding of the
// same entry points is admp_b200/_lib.py (ctypes) + torch.autograd.Function.
//
// Build (where `python -c "import jax; print(jax.ffi.include_dir())"` works):
//   g++ -O2 -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -I../../include \
//       -I/usr/local/cuda/include admp_b200_xla.cc -L../../admp_b200/lib -ladmp_b200 -Wl,-rpath,'$ORIGIN/../../admp_b200/lib' \
//       -o libadmp_b200_xla.so
//
// Conventions: float64 buffers (settings.PRECISION = 'double'; the float32 build swaps F64 for F32), `ctx` = the admp_ctx*
// returned by admp_ctx_create, passed as an int64 attribute (contexts are created / configured from Python through ctypes,
// exactly as admp_b200/_ctx.py does). Optional inputs are passed as zero-sized buffers.
#include <cuda_runtime.h>

#include <cstdint>

#include "admp_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RS32 = ffi::ResultBuffer<ffi::S32>;

static ffi::Error status(int rc) { return rc ? ffi::Error::Internal(admp_last_error()) : ffi::Error::Success(); }
template <typename B> static const void* opt(const B& b) { return b.element_count() ? b.untyped_data() : nullptr; }

// admp/pme.py:58-143, 176-254: get_energy / get_forces / optimize_Uind and every adjoint in one call
static ffi::Error PmeEvalImpl(cudaStream_t stream, int64_t ctx, int32_t flags, int32_t maxiter, double thresh, F64 pos, F64 box, S32 pairs,
                              F64 q_local, F64 u_init, F64 pol, F64 tholes, F64 mscales, F64 pscales, RF64 scalars, RF64 dpos, RF64 dq,
                              RF64 field, RF64 u_out, RF64 dpol, RF64 dtholes, RS32 scf) {
    const bool polz = pol.element_count() != 0;
    if (polz) cudaMemcpyAsync(u_out->untyped_data(), u_init.untyped_data(), u_init.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    return status(admp_pme_eval(reinterpret_cast<admp_ctx*>(ctx), stream, pos.untyped_data(), box.untyped_data(), pairs.typed_data(),
                                pairs.dimensions()[0], q_local.untyped_data(), polz ? u_out->untyped_data() : nullptr, opt(pol),
                                opt(tholes), mscales.untyped_data(), opt(pscales), static_cast<uint32_t>(flags), maxiter, thresh,
                                scalars->typed_data(), dpos->untyped_data(), dq->untyped_data(), polz ? field->untyped_data() : nullptr,
                                polz ? dpol->untyped_data() : nullptr, polz ? dtholes->untyped_data() : nullptr,
                                polz ? scf->typed_data() : nullptr));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(AdmpPmeEval, PmeEvalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx").Attr<int32_t>("flags").Attr<int32_t>("maxiter").Attr<double>("thresh")
                                  .Arg<F64>().Arg<F64>().Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<S32>());

// admp/disp_pme.py:80-279
static ffi::Error DispEvalImpl(cudaStream_t stream, int64_t ctx, int32_t flags, int32_t pmax, F64 pos, F64 box, S32 pairs, F64 c_list,
                               F64 mscales, RF64 scalars, RF64 dpos, RF64 dc) {
    return status(admp_disp_eval(reinterpret_cast<admp_ctx*>(ctx), stream, pos.untyped_data(), box.untyped_data(), pairs.typed_data(),
                                 pairs.dimensions()[0], c_list.untyped_data(), mscales.untyped_data(), pmax, static_cast<uint32_t>(flags),
                                 scalars->typed_data(), dpos->untyped_data(), dc->untyped_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(AdmpDispEval, DispEvalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx").Attr<int32_t>("flags").Attr<int32_t>("pmax")
                                  .Arg<F64>().Arg<F64>().Arg<S32>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>());

// admp/pairwise.py:45-113 with TT_damping_qq_c6_kernel
static ffi::Error TtPairImpl(cudaStream_t stream, int64_t ctx, int32_t flags, F64 pos, F64 box, S32 pairs, F64 mscales, F64 a, F64 b, F64 q,
                             F64 c, RF64 scalars, RF64 dpos, RF64 dparams) {
    return status(admp_tt_pair(reinterpret_cast<admp_ctx*>(ctx), stream, pos.untyped_data(), box.untyped_data(), pairs.typed_data(),
                               pairs.dimensions()[0], mscales.untyped_data(), a.untyped_data(), b.untyped_data(), q.untyped_data(),
                               c.untyped_data(), static_cast<uint32_t>(flags), scalars->typed_data(), dpos->untyped_data(),
                               dparams->untyped_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(AdmpTtPair, TtPairImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx").Attr<int32_t>("flags")
                                  .Arg<F64>().Arg<F64>().Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>());

// jax_md.partition.neighbor_list(..., format=OrderedSparse).allocate / update (examples/water_1024/run_admp.py:109-112).
// NOTE: admp_nblist_build reads the box on the host once (cell-grid geometry) and therefore synchronises the stream.
static ffi::Error NblistImpl(cudaStream_t stream, int64_t ctx, double rc, F64 pos, F64 box, RS32 pairs, RS32 info) {
    return status(admp_nblist_build(reinterpret_cast<admp_ctx*>(ctx), stream, pos.untyped_data(), box.untyped_data(),
                                    static_cast<int>(pos.dimensions()[0]), rc, pairs->typed_data(), pairs->dimensions()[0], info->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(AdmpNblistBuild, NblistImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx").Attr<double>("rc")
                                  .Arg<F64>().Arg<F64>()
                                  .Ret<S32>().Ret<S32>());
