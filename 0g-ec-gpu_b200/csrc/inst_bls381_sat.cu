// inst_bls381_sat.cu -- instantiation unit: the MSM engine over FieldSat<Bls381Fq>.
#include "engine_impl.cuh"

namespace msm {
const FieldOps* field_ops_bls381_sat() {
  static const FieldOps ops = make_field_ops<FieldSat<Bls381Fq>>("bls12-381/sat32");
  return &ops;
}
}  // namespace msm
