// FNLOG{D,I,W,E}: printf-style logging (reference: fyusenet/common/logging.h:30-57).
#pragma once
#include <cstdio>
#ifdef DEBUG
#define FNLOGD(...) do { fprintf(stderr, "[fyn D] " __VA_ARGS__); fputc('\n', stderr); } while (0)
#else
#define FNLOGD(...) do { } while (0)
#endif
#define FNLOGI(...) do { fprintf(stderr, "[fyn I] " __VA_ARGS__); fputc('\n', stderr); } while (0)
#define FNLOGW(...) do { fprintf(stderr, "[fyn W] " __VA_ARGS__); fputc('\n', stderr); } while (0)
#define FNLOGE(...) do { fprintf(stderr, "[fyn E] " __VA_ARGS__); fputc('\n', stderr); } while (0)
