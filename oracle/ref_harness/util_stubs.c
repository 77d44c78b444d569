/*
 * oracle/ref_harness/util_stubs.c -- TEST INFRASTRUCTURE, never shipped.
 *
 * Link-time stand-ins for the util.c symbols that the reference's
 * sim_gravity.c references.  Only the three file-access calls and logmsg do
 * real work; every display entry point aborts, because the harness never
 * enters the display code.  The file-access functions follow the observable
 * behaviour of the reference's util.c:1753-1822 (slurp at most 999,999 bytes
 * into a zeroed 1,000,000 byte buffer, split lines in place at '\n', stop at
 * the first NUL) but go through fopen instead of SDL_RWops.
 */
#include "util.h"

uint32_t sdl_pixel_rgba[MAX_COLOR];

static void display_stub(const char *name)
{
    fprintf(stderr, "ref_harness: display call %s reached in headless oracle\n", name);
    abort();
}

/* pthread_create is renamed to this hook on the command line so that
 * sim_gravity_init() does not start its own workers; the harness starts them. */
int hook_pthread_create(pthread_t *t, const pthread_attr_t *a, void *(*fn)(void *), void *arg)
{
    (void)a; (void)fn; (void)arg;
    *t = 0;
    return 0;
}

void logmsg(char *lvl, const char *func, char *fmt, ...)
{
    if (getenv("REF_HARNESS_VERBOSE") == NULL) return;
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "%s %s: ", lvl, func);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

void printmsg(char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

typedef struct { char *buff; size_t offset; } stub_file_t;

file_t *open_file(char *pathname)
{
    enum { CAP = 1000000 };
    FILE *fp = fopen(pathname, "r");
    if (fp == NULL) return NULL;
    char *buff = calloc(1, CAP);
    size_t len = fread(buff, 1, CAP - 1, fp);
    fclose(fp);
    if (len == 0) { free(buff); return NULL; }
    stub_file_t *f = calloc(1, sizeof(*f));
    f->buff = buff;
    return f;
}

char *read_file_line(file_t *file)
{
    stub_file_t *f = file;
    if (f->buff[f->offset] == '\0') return NULL;
    char *s = f->buff + f->offset;
    char *nl = strchr(s, '\n');
    if (nl != NULL) { *nl = '\0'; f->offset = (size_t)(nl - f->buff) + 1; }
    else            { f->offset = (size_t)(s - f->buff) + strlen(s); }
    return s;
}

void close_file(file_t *file)
{
    stub_file_t *f = file;
    if (f == NULL) return;
    free(f->buff);
    free(f);
}

int32_t list_files(char *location, int32_t *max, char ***pathnames)
{ (void)location; *max = 0; *pathnames = NULL; return -1; }
void list_files_free(int32_t max, char **pathnames) { (void)max; (void)pathnames; }

uint64_t microsec_timer(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000 + ts.tv_nsec / 1000;
}

char *dur2str(char *str, int64_t duration) { sprintf(str, "%lld", (long long)duration); return str; }

void sdl_event_init(void) { display_stub(__func__); }
void sdl_event_register(int32_t e, int32_t t, SDL_Rect *p) { (void)e; (void)t; (void)p; display_stub(__func__); }
sdl_event_t *sdl_poll_event(void) { display_stub(__func__); return NULL; }
void sdl_render_text_ex(SDL_Rect *pane, int32_t row, int32_t col, char *str, int32_t event,
                        int32_t field_cols, bool center, int32_t font_id)
{ (void)pane; (void)row; (void)col; (void)str; (void)event; (void)field_cols; (void)center; (void)font_id; display_stub(__func__); }
void sdl_render_text_font0(SDL_Rect *pane, int32_t row, int32_t col, char *str, int32_t event)
{ (void)pane; (void)row; (void)col; (void)str; (void)event; display_stub(__func__); }
void sdl_render_text_font1(SDL_Rect *pane, int32_t row, int32_t col, char *str, int32_t event)
{ (void)pane; (void)row; (void)col; (void)str; (void)event; display_stub(__func__); }
void sdl_render_rect(SDL_Rect *rect, int32_t lw, uint32_t rgba) { (void)rect; (void)lw; (void)rgba; display_stub(__func__); }
SDL_Texture *sdl_create_filled_circle_texture(int32_t r, uint32_t rgba) { (void)r; (void)rgba; display_stub(__func__); return NULL; }
void sdl_render_circle(int32_t x, int32_t y, SDL_Texture *t) { (void)x; (void)y; (void)t; display_stub(__func__); }
void sdl_display_text(char *text) { (void)text; display_stub(__func__); }
void sdl_display_choose_from_list(char *t, char **c, int32_t n, int32_t *s) { (void)t; (void)c; (void)n; (void)s; display_stub(__func__); }
void sdl_display_error(char *a, char *b, char *c) { (void)a; (void)b; (void)c; display_stub(__func__); }

int SDL_SetRenderDrawColor(SDL_Renderer *r, uint8_t cr, uint8_t cg, uint8_t cb, uint8_t ca)
{ (void)r; (void)cr; (void)cg; (void)cb; (void)ca; display_stub(__func__); return 0; }
int SDL_RenderClear(SDL_Renderer *r) { (void)r; display_stub(__func__); return 0; }
int SDL_RenderDrawLines(SDL_Renderer *r, const SDL_Point *p, int n) { (void)r; (void)p; (void)n; display_stub(__func__); return 0; }
int SDL_RenderFillRect(SDL_Renderer *r, const SDL_Rect *q) { (void)r; (void)q; display_stub(__func__); return 0; }
void SDL_RenderPresent(SDL_Renderer *r) { (void)r; display_stub(__func__); }
