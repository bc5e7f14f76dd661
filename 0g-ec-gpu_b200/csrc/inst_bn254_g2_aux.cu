// inst_bn254_g2_aux.cu -- second instantiation unit of FieldExt2Lazy<Bn254Fq>: EC-FFT, synthetic inputs, affine conversion,
// point sums, per-primitive test kernels (see inst_bn254_g2.cu).
#include "engine_impl.cuh"

namespace msm {
void fill_field_ops_aux_bn254_g2(FieldOps& o) { fill_field_ops_aux<FieldExt2Lazy<Bn254Fq>>(o); }
}  // namespace msm
