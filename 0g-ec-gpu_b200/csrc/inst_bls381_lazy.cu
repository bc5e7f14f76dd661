// inst_bls381_lazy.cu -- instantiation unit: the MSM engine over FieldSatLazy<Bls381Fq>.
#include "engine_impl.cuh"

namespace msm {
const FieldOps* field_ops_bls381_lazy() {
  static const FieldOps ops = make_field_ops<FieldSatLazy<Bls381Fq>>("bls12-381/sat32-lazy");
  return &ops;
}
}  // namespace msm
