#include "opencv2/cvstub.hpp"
