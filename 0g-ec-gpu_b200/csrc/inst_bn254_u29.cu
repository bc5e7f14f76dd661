// inst_bn254_u29.cu -- instantiation unit: the MSM engine over FieldU29<Bn254U29>.
#include "engine_impl.cuh"

namespace msm {
const FieldOps* field_ops_bn254_u29() {
  static const FieldOps ops = make_field_ops<FieldU29<Bn254U29>>("bn254/u29");
  return &ops;
}
}  // namespace msm
