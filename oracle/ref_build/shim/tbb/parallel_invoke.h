#pragma once
#include "tbb.h"
