/* stub (written for this repo): the athread runtime does not exist on x86.  The MPE-side SERIAL routines of the *_sw64.c
 * files never spawn; the spawn/join entry points exist only so the files compile and link, and trap if reached. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#define SLAVE_FUN(name) void slave_##name
static inline int athread_idle(void) { return 1; }
static inline int athread_init(void) { return 0; }
static inline int athread_join(void) { return 0; }
#define athread_spawn(fn, arg) do { fprintf(stderr, "oracle/_ref: athread_spawn reached\n"); abort(); } while (0)
