// sphere.h — same name as the reference header; everything lives in rtb_host.hpp
#include "../rtb_host.hpp"
