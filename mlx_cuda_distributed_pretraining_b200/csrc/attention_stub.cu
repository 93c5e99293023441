// temporary: replaced by attention_fwd.cu / attention_bwd.cu
#include "host.h"
namespace b200 {
int attn_fwd(const void*, const void*, const void*, void*, float*, int, int, int, int, int, float, int, cudaStream_t) {
  set_error("attn_fwd: not built"); return B200_ERR_ARG; }
size_t attn_bwd_workspace_bytes(int, int, int, int, int) { return 0; }
int attn_bwd(const void*, const void*, const void*, const void*, const void*, const float*, void*, void*, void*, int, int, int, int, int, float, int, void*, size_t, cudaStream_t) {
  set_error("attn_bwd: not built"); return B200_ERR_ARG; }
}
