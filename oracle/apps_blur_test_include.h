// Indirection so the include path of the reference source is visible in one place.
#include "/root/reference/apps/blur/test.cpp"
