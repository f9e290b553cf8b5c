#include "highgui/highgui.hpp"
