/* cvshim: display / image file functions — declared so that the reference's headers compile; never on the depth path */
#ifndef CVSHIM_HIGHGUI_HPP
#define CVSHIM_HIGHGUI_HPP
#include "../core/core.hpp"
namespace cv {
inline void namedWindow(const std::string&, int = 1) { shim_abort("namedWindow"); }
inline void resizeWindow(const std::string&, int, int) { shim_abort("resizeWindow"); }
inline void imshow(const std::string&, const Mat&) { shim_abort("imshow"); }
inline int waitKey(int = 0) { shim_abort("waitKey"); }
inline void destroyAllWindows() { shim_abort("destroyAllWindows"); }
inline Mat imread(const std::string&, int = 1) { shim_abort("imread"); }
inline bool imwrite(const std::string&, const Mat&, const std::vector<int>& = std::vector<int>()) { shim_abort("imwrite"); }
}
#endif
