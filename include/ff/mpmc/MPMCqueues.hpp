// FastFlow-compatible runtime: MPMC_Ptr_Queue lives in ff/ff.hpp
#pragma once
#include "../ff.hpp"
