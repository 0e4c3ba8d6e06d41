/* config.h for the oracle/_ref build of the unmodified reference (test infrastructure). */
#define ENABLE_ENCODER 1
#define ENABLE_MOTION_REF 1
#ifndef DISABLE_ORC
#define DISABLE_ORC 1
#endif
#define VERSION "1.0.11.1-oracle"
#define HAVE_THREAD_PTHREAD 1
