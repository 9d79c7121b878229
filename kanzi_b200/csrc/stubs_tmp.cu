// temporary stubs (replaced as kernels land)
#include "kzg_common.cuh"
#include "kzg_entropy.cuh"
#include "kzg_xf_kernels.cuh"
#include "kzg_stages.cuh"
void kzg_bwt_scratch(i32, bool, size_t*, size_t*) {}
void kzg_rolz_scratch(i32, bool, size_t*, size_t*, size_t*) {}
void kzg_small_scratch(int, i32, bool, size_t*, size_t*) {}
int kzg_zrlt_launch(cudaStream_t, bool, KzgBlock*, int, const KzgXfParams&) { return -KZG_ERR_INVALID_CODEC; }
int kzg_sbrt_launch(cudaStream_t, bool, int, KzgBlock*, int, const KzgXfParams&) { return -KZG_ERR_INVALID_CODEC; }
int kzg_srt_launch(cudaStream_t, bool, KzgBlock*, int, const KzgXfParams&) { return -KZG_ERR_INVALID_CODEC; }
int kzg_bwtblock_launch(cudaStream_t, bool, KzgBlock*, int, const KzgXfParams&, i32) { return -KZG_ERR_INVALID_CODEC; }
int kzg_rolz_launch(cudaStream_t, bool, KzgBlock*, int, const KzgXfParams&, i32) { return -KZG_ERR_INVALID_CODEC; }
int kzg_bwt_raw(cudaStream_t, bool, const u8*, i32, u8*, i32*) { return -KZG_ERR_INVALID_CODEC; }
int kzg_fpaq_encode_launch(cudaStream_t, const KzgBlock*, int, const KzgEntParams&) { return -KZG_ERR_INVALID_CODEC; }
int kzg_fpaq_decode_launch(cudaStream_t, KzgBlock*, int, const KzgEntParams&) { return -KZG_ERR_INVALID_CODEC; }
