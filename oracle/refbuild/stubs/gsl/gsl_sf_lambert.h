// Stand-in: see gsl_math.h in this directory.
#pragma once
#include "gsl_math.h"
