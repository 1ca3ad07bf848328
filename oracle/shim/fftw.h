#include "fftw_types.h"
