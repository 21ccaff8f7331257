/* Symbol visibility macros (the reference generates these with CMake GenerateExportHeader). */
#ifndef SPFFT_EXPORT_H
#define SPFFT_EXPORT_H
#if defined(__GNUC__) || defined(__clang__)
#define SPFFT_EXPORT __attribute__((visibility("default")))
#define SPFFT_NO_EXPORT __attribute__((visibility("hidden")))
#else
#define SPFFT_EXPORT
#define SPFFT_NO_EXPORT
#endif
#endif
