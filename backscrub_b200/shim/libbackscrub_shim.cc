// backscrub_b200/shim/libbackscrub_shim.cc — C++-linkage definitions of the reference's four entry points
// (lib/libbackscrub.h:13-39) over the C ABI of libbackscrub_b200.so.  Compile this file (it needs OpenCV *core*
// headers only) into the reference's `backscrub` library target in place of lib/libbackscrub.cc; app/deepseg.cc
// then links unchanged whichever lib/libbackscrub.h its include path finds.
#include <opencv2/core/core.hpp>

#include <string>

#include "backscrub_b200.h"

const char *bs_tensorflow_version(void) { return bsb_version(); }

void *bs_maskgen_new(const std::string& modelname, size_t threads, size_t width, size_t height,
                     void (*ondebug)(void *ctx, const char *msg), void (*onprep)(void *ctx),
                     void (*oninfer)(void *ctx), void (*onmask)(void *ctx), void *caller_ctx) {
	return bsb_maskgen_new(modelname.c_str(), threads, width, height, ondebug, onprep, oninfer, onmask, caller_ctx);
}

void bs_maskgen_delete(void *context) { bsb_maskgen_delete(static_cast<bsb_ctx *>(context)); }

bool bs_maskgen_process(void *context, cv::Mat& frame, cv::Mat &mask) {
	if (!context || frame.empty() || frame.type() != CV_8UC3)
		return false;
	bsb_ctx *ctx = static_cast<bsb_ctx *>(context);
	int w = 0, h = 0;
	if (!bsb_frame_size(ctx, &w, &h))
		return false;
	// The reference crops with frame(ctx.roidim): a frame smaller than the context size makes cv::Mat::operator()
	// throw; a larger one is read through its top-left width x height window (lib/libbackscrub.cc:285).  No
	// exception crosses this boundary, so a short frame is an error return.
	if (frame.cols < w || frame.rows < h)
		return false;
	const uint8_t *mptr = nullptr;
	size_t mpitch = 0;
	if (!bsb_maskgen_process(ctx, frame.data, frame.step, &mptr, &mpitch))
		return false;
	// header over context-owned storage, like `mask = ctx.mask` (lib/libbackscrub.cc:374): always width x height
	mask = cv::Mat(h, w, CV_8UC1, const_cast<uint8_t *>(mptr), mpitch);
	return true;
}
