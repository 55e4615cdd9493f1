#pragma once
#include "pb_shim.h"
