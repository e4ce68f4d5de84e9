// include path of the reference kept for its callers; the structs live in utils/b200_params.h
#pragma once
#include "../../utils/b200_params.h"
