// Drop-in for the reference's src/models/common_params.h (an empty file there, included by basemodel.h:7): kept so that the include resolves.
#pragma once
