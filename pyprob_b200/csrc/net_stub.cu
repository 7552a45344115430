// TEMPORARY bring-up stubs (replaced by net_*.cu): keep the ABI complete while the first kernels are validated.
#include "common.cuh"
extern "C" {
#define STUB(name, ...) int name(__VA_ARGS__) { ppb_set_error(#name ": not implemented yet"); return PPB_ENOTSUP; }
STUB(ppb_net_create, ppb_net**, const ppb_net_desc*)
STUB(ppb_net_set_tables, ppb_net*, const ppb_addr_desc*, int32_t, const int64_t*, int32_t, int64_t)
STUB(ppb_net_destroy, ppb_net*)
int64_t ppb_ic_workspace_bytes(const ppb_net*, int32_t, int32_t, int32_t, int32_t) { return -1; }
STUB(ppb_ic_loss_forward, ppb_net*, const float*, const ppb_batch*, void*, int64_t, int, float*, int32_t*, float*, void*)
STUB(ppb_ic_loss_backward, ppb_net*, const float*, float*, const ppb_batch*, void*, int64_t, int, float, void*)
STUB(ppb_adam_step, float*, const float*, float*, float*, int64_t, float, float, float, float, float, int64_t, float, void*)
STUB(ppb_ic_infer_step, ppb_net*, const float*, const float*, int, int32_t, const float*, int32_t, const float*, int, const float*, int, float*, float*, float*, int64_t, void*, int64_t, int, void*)
STUB(ppb_ic_embed_observe, ppb_net*, const float*, const float*, float*, int64_t, void*, int64_t, void*)
int64_t ppb_ic_infer_workspace_bytes(const ppb_net*, int64_t) { return -1; }
STUB(ppb_ic_train_step_host, ppb_net*, float*, float*, float*, float*, int64_t, const void*, int64_t, void*, void*, int64_t, int, float, float, float, float, float, int64_t, float*, int32_t*, void*)
}
