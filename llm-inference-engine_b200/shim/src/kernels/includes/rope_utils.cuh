// The reference's rope_utils.cuh holds device helpers of its own kernels; nothing of it is part of the host API.
#pragma once
