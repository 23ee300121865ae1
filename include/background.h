/* include/background.h — the reference's app/background.h:14-23 interface, provided by the B200 library.
 *
 * Same three C++-linkage functions with the same signatures and return values; they are defined by
 * backscrub_b200/shim/background_shim.cc (compiled by the integrator in place of app/background.cc).  Decoding
 * stays on the host (cv::VideoCapture / cv::imread); the reader thread's pacing / looping / frame counting live in
 * the library's provider object (bsb_background_*, include/backscrub_b200.h) and grab_background()'s per-frame
 * cv::resize runs on the GPU.
 */
#ifndef _BACKGROUND_H_
#define _BACKGROUND_H_

#include <opencv2/core/mat.hpp>

#include <memory>
#include <string>

struct background_t;

// Load a background (image or video file, stream URL).  nullptr on error; the handle cleans up after itself.
std::shared_ptr<background_t> load_background(const std::string& path, int debug);

// Latest background frame resized to width x height.  Returns the frame number (1 for a still image; a looping
// video wraps to 0) or -1 on error.
int grab_background(std::shared_ptr<background_t> handle, int width, int height, cv::Mat &out);

// Copy of the current thumbnail (empty until one exists).  <0 on error, 0 on success.
int grab_thumbnail(std::shared_ptr<background_t> handle, cv::Mat &out);

#endif
