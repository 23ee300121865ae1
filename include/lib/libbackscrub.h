/* include/lib/libbackscrub.h — the reference's lib/libbackscrub.h interface (lib/libbackscrub.h:13-39), provided by
 * the B200 library.
 *
 * app/deepseg.cc:24 says `#include "lib/libbackscrub.h"` and the reference's CMakeLists.txt:72 puts the source root
 * first on the include path, so inside the reference tree the reference's OWN header is the one that is found; with
 * `-I <this repo>/include` ahead of it, this one is.  Either way the declarations are the same four C++-linkage
 * functions, and they are DEFINED (C++ linkage, std::string / cv::Mat& signatures) by
 * backscrub_b200/shim/libbackscrub_shim.cc, which the integrator compiles into the `backscrub` target instead of
 * lib/libbackscrub.cc + lib/transpose_conv_bias.cc + the TFLite subtree.  app/deepseg.cc is not touched.
 */
#ifndef _LIBBACKSCRUB_H
#define _LIBBACKSCRUB_H

// for cv::Mat
#include <opencv2/core/core.hpp>

#include <string>

// Name of the inference runtime (the reference returns TFLITE_VERSION_STRING; printed by deepseg.cc:351)
extern const char *bs_tensorflow_version(void);

// New opaque mask-generation context, or nullptr after reporting through ondebug (stderr when ondebug is null).
// Callbacks are optional; onprep / oninfer / onmask fire in this order once per bs_maskgen_process call.
extern void *bs_maskgen_new(
	const std::string& modelname,
	size_t threads,
	size_t width,
	size_t height,
	void (*ondebug)(void *ctx, const char *msg),
	void (*onprep)(void *ctx),
	void (*oninfer)(void *ctx),
	void (*onmask)(void *ctx),
	void *caller_ctx
);

// Delete the context (nullptr-safe)
extern void bs_maskgen_delete(void *context);

// One BGR frame in, mask out: `mask` becomes a header over context-owned storage, valid until the next call
extern bool bs_maskgen_process(void *context, cv::Mat& frame, cv::Mat &mask);

#endif
