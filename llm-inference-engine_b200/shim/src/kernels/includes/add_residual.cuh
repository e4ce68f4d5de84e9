// Drop-in for the reference's src/kernels/includes/add_residual.cuh: the launcher(s) declared there are provided, with the
// same signature, by b200_launchers.h on top of libb200llm.so.
#pragma once
#include "b200_launchers.h"
