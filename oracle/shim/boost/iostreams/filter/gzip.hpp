#pragma once
#include "../filtering_stream.hpp"
