// render_stub.cu — placeholder until restir.cu / bounce.cu / northstar.cu land
#include "kernels.h"
namespace snk {
int launch_reference_frame(SnkPt* ctx, const SnkRtSettings&, cudaStream_t) { ctx->err = "reference frame: not built yet"; return SNKPT_E_STATE; }
int launch_northstar_frame(SnkPt* ctx, const SnkRtSettings&, cudaStream_t) { ctx->err = "northstar frame: not built yet"; return SNKPT_E_STATE; }
}
