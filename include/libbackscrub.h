/* include/libbackscrub.h — drop-in replacement for the reference's lib/libbackscrub.h.
 *
 * Same four functions, same signatures (lib/libbackscrub.h:13-39), so app/deepseg.cc
 * compiles against this header unchanged and links with -lbackscrub_b200 instead of the
 * TFLite/XNNPACK-based libbackscrub.  Header-only: every function is a thin adapter from
 * cv::Mat / std::string to the C ABI in backscrub_b200.h.  It needs OpenCV *core* headers
 * only (cv::Mat), exactly like the header it replaces.
 *
 * Mask aliasing follows the reference: `mask` becomes a cv::Mat header over storage owned by
 * the context, valid until the next bs_maskgen_process() on that context
 * (lib/libbackscrub.cc:374 does `mask = ctx.mask`).
 */
#ifndef _LIBBACKSCRUB_H
#define _LIBBACKSCRUB_H

#include <opencv2/core/core.hpp>

#include <string>

#include "backscrub_b200.h"

// Get the inference runtime's version string (the reference returns TFLITE_VERSION_STRING)
static inline const char *bs_tensorflow_version(void) { return bsb_version(); }

// Return a new (opaque) mask generation context
static inline void *bs_maskgen_new(
	const std::string& modelname,
	size_t threads,
	size_t width,
	size_t height,
	void (*ondebug)(void *ctx, const char *msg),
	void (*onprep)(void *ctx),
	void (*oninfer)(void *ctx),
	void (*onmask)(void *ctx),
	void *caller_ctx
) {
	return bsb_maskgen_new(modelname.c_str(), threads, width, height, ondebug, onprep, oninfer, onmask, caller_ctx);
}

// Delete the mask generation context
static inline void bs_maskgen_delete(void *context) { bsb_maskgen_delete(static_cast<bsb_ctx *>(context)); }

// Process a video frame into a mask
static inline bool bs_maskgen_process(void *context, cv::Mat& frame, cv::Mat &mask) {
	if (!context || frame.type() != CV_8UC3)
		return false;
	const uint8_t *mptr = nullptr;
	size_t mpitch = 0;
	if (!bsb_maskgen_process(static_cast<bsb_ctx *>(context), frame.data, frame.step, &mptr, &mpitch))
		return false;
	// header over context-owned storage, like `mask = ctx.mask` in the reference
	mask = cv::Mat(frame.rows, frame.cols, CV_8UC1, const_cast<uint8_t *>(mptr), mpitch);
	return true;
}

#endif
