// include path of the reference kept for its callers; the aliases live in b200_params.h
#pragma once
#include "b200_params.h"
