// Builds the C++ host layer (calibr8_b200/host) into libc8b200.so.
#include "../host/host.cu"
#include "../host/vfm_host.cu"
#include "../host/objective.cu"
