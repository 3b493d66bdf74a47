#pragma once
#include "cpubuffer.h"
