// Oracle shim: gflags is absent; nothing of it is used by src/ORBextractor.cpp.
