/* Stand-in for the CMake GenerateExportHeader output the reference build produces. */
#ifndef SPFFT_EXPORT_H
#define SPFFT_EXPORT_H
#define SPFFT_EXPORT __attribute__((visibility("default")))
#define SPFFT_NO_EXPORT __attribute__((visibility("hidden")))
#define SPFFT_DEPRECATED __attribute__((__deprecated__))
#endif
