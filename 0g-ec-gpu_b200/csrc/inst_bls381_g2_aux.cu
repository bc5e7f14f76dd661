// inst_bls381_g2_aux.cu -- second instantiation unit of FieldExt2Lazy<Bls381Fq>: EC-FFT, synthetic inputs, affine conversion,
// point sums, per-primitive test kernels (see inst_bls381_g2.cu).
#include "engine_impl.cuh"

namespace msm {
void fill_field_ops_aux_bls381_g2(FieldOps& o) { fill_field_ops_aux<FieldExt2Lazy<Bls381Fq>>(o); }
}  // namespace msm
