// inst_bn254_lazy.cu -- instantiation unit: the MSM engine over FieldSatLazy<Bn254Fq>.
#include "engine_impl.cuh"

namespace msm {
const FieldOps* field_ops_bn254_lazy() {
  static const FieldOps ops = make_field_ops<FieldSatLazy<Bn254Fq>>("bn254/sat32-lazy");
  return &ops;
}
}  // namespace msm
