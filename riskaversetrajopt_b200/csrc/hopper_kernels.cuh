#pragma once
#include "saa_common.cuh"
