// TEST INFRASTRUCTURE ONLY (oracle/_ref): see NvInfer.h in this directory.
#pragma once
#include "NvInfer.h"
