// inst_bn254_g2.cu -- instantiation unit: the MSM engine over FieldExt2Lazy<Bn254Fq> (G2).
#include "engine_impl.cuh"

namespace msm {
const FieldOps* field_ops_bn254_g2() {
  static const FieldOps ops = make_field_ops<FieldExt2Lazy<Bn254Fq>>("bn254-g2/fq2-sat32-lazy");
  return &ops;
}
}  // namespace msm
