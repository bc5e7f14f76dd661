// inst_bn254_sat.cu -- instantiation unit: the MSM engine over FieldSat<Bn254Fq>.
#include "engine_impl.cuh"

namespace msm {
const FieldOps* field_ops_bn254_sat() {
  static const FieldOps ops = make_field_ops<FieldSat<Bn254Fq>>("bn254/sat32");
  return &ops;
}
}  // namespace msm
