// backscrub_b200/shim/background_shim.cc — the reference's background provider interface (app/background.h:14-23)
// over the library's provider object.  Compile this file in place of app/background.cc (it needs OpenCV videoio +
// imgcodecs for decoding, like the file it replaces).  What stays here is I/O: opening the media, the "can I read
// two frames" video probe, and handing decoded frames to the library; the reader thread with its real-time pacing,
// looping and frame counting, and the per-grab resize (on the GPU) are in libbackscrub_b200.so (csrc/frontend.cu).
#include <opencv2/core/mat.hpp>
#include <opencv2/imgcodecs.hpp>
#include <opencv2/videoio.hpp>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

#include "background.h"
#include "backscrub_b200.h"

struct background_t {
	int debug = 0;
	bool video = false;
	cv::VideoCapture cap;
	cv::Mat grab;                 // the frame the decoder last produced (owned here, read by the library's thread)
	bsb_background *provider = nullptr;
	int device = 0;
};

static int shim_read(void *user, const uint8_t **data, int *w, int *h, size_t *pitch) {
	background_t *b = static_cast<background_t *>(user);
	if (!b->cap.read(b->grab) || b->grab.empty() || b->grab.type() != CV_8UC3)
		return 0;
	*data = b->grab.data; *w = b->grab.cols; *h = b->grab.rows; *pitch = b->grab.step;
	return 1;
}

static int shim_rewind(void *user) {
	background_t *b = static_cast<background_t *>(user);
	return b->cap.set(cv::CAP_PROP_POS_FRAMES, 0) ? 1 : 0;
}

static void drop_background(background_t *b) {
	if (!b)
		return;
	bsb_background_delete(b->provider);     // stops and joins the reader thread first
	if (b->video)
		b->cap.release();
	delete b;
}

std::shared_ptr<background_t> load_background(const std::string& path, int debug) {
	auto b = std::shared_ptr<background_t>(new background_t, drop_background);
	try {
		b->debug = debug;
		const char *dev = std::getenv("BSB_DEVICE");
		b->device = dev ? std::atoi(dev) : 0;
		b->cap.open(path, cv::CAP_ANY);
		if (!b->cap.isOpened()) {
			if (debug) fprintf(stderr, "background: cap cannot open: %s\n", path.c_str());
			return nullptr;
		}
		b->cap.set(cv::CAP_PROP_CONVERT_RGB, true);
		const double fps = b->cap.get(cv::CAP_PROP_FPS);
		// two readable frames => video; else try it as an image; else unusable (app/background.cc:141-162)
		cv::Mat first;
		if (b->cap.read(first) && b->cap.read(first)) {
			const int start = b->cap.set(cv::CAP_PROP_POS_FRAMES, 0) ? 0 : 2;
			b->video = true;
			b->provider = bsb_background_new_video(b->device, fps, start, shim_read, shim_rewind, b.get(),
			                                       first.data, first.cols, first.rows, first.step, debug);
		} else {
			b->cap.release();
			cv::Mat img = cv::imread(path);
			if (img.empty()) {
				if (debug) fprintf(stderr, "background: imread cannot open: %s\n", path.c_str());
				return nullptr;
			}
			b->provider = bsb_background_new_still(b->device, img.data, img.cols, img.rows, img.step, debug);
		}
		if (!b->provider)
			return nullptr;
		if (debug)
			fprintf(stderr, "background properties:\n\tvid: %s\n\tfps: %f\n", b->video ? "yes" : "no", fps);
	} catch (...) {
		if (debug) fprintf(stderr, "background: exception while loading\n");
		return nullptr;
	}
	return b;
}

int grab_background(std::shared_ptr<background_t> b, int width, int height, cv::Mat &out) {
	if (!b || !b->provider)
		return -1;
	out.create(height, width, CV_8UC3);
	return bsb_background_grab(b->provider, width, height, out.data, out.step);
}

int grab_thumbnail(std::shared_ptr<background_t> b, cv::Mat &out) {
	if (!b || !b->provider)
		return -1;
	int w = 0, h = 0;
	if (bsb_background_thumbnail(b->provider, nullptr, 0, &w, &h) < 0)
		return -1;
	if (w <= 0 || h <= 0) {
		out = cv::Mat();                      // no thumbnail yet: the reference clones an empty Mat
		return 0;
	}
	out.create(h, w, CV_8UC3);
	return bsb_background_thumbnail(b->provider, out.data, (size_t)w * h * 3, &w, &h);
}
