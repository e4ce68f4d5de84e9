// include path of the reference kept for its callers; the type itself lives in b200_model_types.h
#pragma once
#include "b200_model_types.h"
