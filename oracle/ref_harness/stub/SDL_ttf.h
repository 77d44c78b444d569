/* Minimal stand-in for <SDL_ttf.h>; see stub/SDL.h. */
#ifndef ORACLE_STUB_SDL_TTF_H
#define ORACLE_STUB_SDL_TTF_H
typedef struct TTF_Font TTF_Font;
#endif
