// inst_bls381_g2.cu -- instantiation unit: the MSM path over FieldExt2Lazy<Bls381Fq> (G2).
// The EC-FFT, helper and test kernels of the same field class are instantiated in inst_bls381_g2_aux.cu (the Fq2 units
// are the slowest to compile; two units build in parallel).
#include "engine_impl.cuh"

namespace msm {
const FieldOps* field_ops_bls381_g2() {
  static const FieldOps ops = [] {
    FieldOps o;
    fill_field_ops_msm<FieldExt2Lazy<Bls381Fq>>(o, "bls12-381-g2/fq2-sat32-lazy");
    fill_field_ops_aux_bls381_g2(o);
    return o;
  }();
  return &ops;
}
}  // namespace msm
