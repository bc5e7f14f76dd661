// inst_bls381_g2.cu -- instantiation unit: the MSM engine over FieldExt2Lazy<Bls381Fq> (G2).
#include "engine_impl.cuh"

namespace msm {
const FieldOps* field_ops_bls381_g2() {
  static const FieldOps ops = make_field_ops<FieldExt2Lazy<Bls381Fq>>("bls12-381-g2/fq2-sat32-lazy");
  return &ops;
}
}  // namespace msm
