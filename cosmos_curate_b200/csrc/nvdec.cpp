// NVDEC session management (placeholder until the cuvid path lands in this file).
#include "common.h"
extern "C" void cb_nvdec_release(cb_ctx* ctx) { (void)ctx; }
