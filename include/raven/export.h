// raven-b200: stands in for the CMake-generated export header of RavenLib
// (RavenLib/RavenLib.cmake:29 generates raven/export.h).
#ifndef RAVEN_EXPORT_H_
#define RAVEN_EXPORT_H_
#ifndef RAVEN_EXPORT
#define RAVEN_EXPORT __attribute__((visibility("default")))
#endif
#ifndef RAVEN_NO_EXPORT
#define RAVEN_NO_EXPORT __attribute__((visibility("hidden")))
#endif
#endif  // RAVEN_EXPORT_H_
