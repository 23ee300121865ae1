/* include/background.h — the reference's app/background.h:14-23 API with grab_background()'s
 * per-frame cv::resize moved to the GPU.
 *
 * Decoding (cv::VideoCapture / cv::imread, the reader thread and its pacing,
 * app/background.cc:13-176) is I/O and stays in the application: keep the reference's
 * app/background.cc for load_background()/grab_thumbnail() and replace only the body of
 * grab_background() with the adapter below, or call bsb_set_background() directly whenever a
 * new decoded frame is available and let bsb_composite() blend against the resident copy.
 */
#ifndef _BACKGROUND_B200_H_
#define _BACKGROUND_B200_H_

#include <opencv2/core/mat.hpp>

#include "backscrub_b200.h"

// app/background.cc:178-194: cv::resize(raw, out, cv::Size(width, height)) — on the GPU.
// `raw` is the decoded background frame (CV_8UC3); returns 0 on success, -1 on error.
static inline int bsb_grab_background(bsb_ctx *ctx, const cv::Mat &raw, int width, int height, cv::Mat &out) {
	if (!ctx || raw.empty() || raw.type() != CV_8UC3)
		return -1;
	if (!bsb_set_background(ctx, raw.data, raw.cols, raw.rows, raw.step))
		return -1;
	out.create(height, width, CV_8UC3);
	return bsb_get_background(ctx, out.data, out.step) ? 0 : -1;
}

#endif
