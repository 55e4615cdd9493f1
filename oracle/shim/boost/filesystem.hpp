// Stand-in: the reference includes boost/filesystem.hpp (src/WEPP/dataset.hpp:4, util.cpp:2)
// but the placement translation units use nothing from it.
#pragma once
