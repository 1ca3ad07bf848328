#include "rfftw_naive.h"
