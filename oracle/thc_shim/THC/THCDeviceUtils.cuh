// see THC.h in this directory
#pragma once
#include <ATen/cuda/DeviceUtils.cuh>
