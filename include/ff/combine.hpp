// FastFlow-compatible runtime: everything lives in ff/ff.hpp
#pragma once
#include "ff.hpp"
