// Stand-in for the protoc-generated header the reference's MAT header includes
// (src/mutation_annotated_tree.hpp:21).  The shim build never compiles the protobuf loader
// (src/mutation_annotated_tree.cpp:522-612); oracle/ref_driver.cpp injects trees directly.
#pragma once
