// see gpulayerbuilder.h (all GPU builders live there)
#pragma once
#include "gpulayerbuilder.h"
