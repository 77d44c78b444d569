/* Minimal stand-in for <SDL.h>, used ONLY to compile the reference's
 * sim_gravity.c translation unit headless inside oracle/_ref (test
 * infrastructure, never shipped).  It declares just enough types and
 * prototypes for the display code to parse; none of it is ever executed by
 * the harness.  SDL2 itself is not installed in this image. */
#ifndef ORACLE_STUB_SDL_H
#define ORACLE_STUB_SDL_H
#include <stdint.h>
typedef struct SDL_Window SDL_Window;
typedef struct SDL_Renderer SDL_Renderer;
typedef struct SDL_Texture SDL_Texture;
typedef struct SDL_Rect { int x, y, w, h; } SDL_Rect;
typedef struct SDL_Point { int x, y; } SDL_Point;
#define SDL_ALPHA_OPAQUE 255
int SDL_SetRenderDrawColor(SDL_Renderer *r, uint8_t cr, uint8_t cg, uint8_t cb, uint8_t ca);
int SDL_RenderClear(SDL_Renderer *r);
int SDL_RenderDrawLines(SDL_Renderer *r, const SDL_Point *pts, int count);
int SDL_RenderFillRect(SDL_Renderer *r, const SDL_Rect *rect);
void SDL_RenderPresent(SDL_Renderer *r);
#endif
