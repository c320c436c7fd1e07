// TEST INFRASTRUCTURE ONLY (oracle/_ref): forwards to the CPU mock of the OpenCV subset the reference uses.
#pragma once
#include "w2x_cvshim.hpp"
