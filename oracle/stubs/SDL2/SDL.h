/* Minimal stand-in for SDL2's header: just the names the reference's main()
 * and output thread mention, so that the reference translation unit compiles.
 * None of these functions is ever called by the oracle (the player's main is
 * renamed and dead).  Test infrastructure only. */
#ifndef FM_ORACLE_SDL_STUB_H
#define FM_ORACLE_SDL_STUB_H
#include <stdint.h>
typedef uint32_t SDL_AudioDeviceID;
typedef uint16_t SDL_AudioFormat;
typedef void (*SDL_AudioCallback)(void *userdata, uint8_t *stream, int len);
typedef struct SDL_AudioSpec {
    int freq;
    SDL_AudioFormat format;
    uint8_t channels;
    uint8_t silence;
    uint16_t samples;
    uint16_t padding;
    uint32_t size;
    SDL_AudioCallback callback;
    void *userdata;
} SDL_AudioSpec;
#define AUDIO_S16LSB 0x8010
#define AUDIO_S16SYS AUDIO_S16LSB
#define AUDIO_S16 AUDIO_S16LSB
#define SDL_INIT_AUDIO 0x00000010u
#define SDL_AUDIO_ALLOW_ANY_CHANGE 0x0f
int SDL_Init(uint32_t flags);
void SDL_Quit(void);
const char *SDL_GetError(void);
const char *SDL_GetCurrentAudioDriver(void);
SDL_AudioDeviceID SDL_OpenAudioDevice(const char *device, int iscapture, const SDL_AudioSpec *desired,
                                      SDL_AudioSpec *obtained, int allowed_changes);
void SDL_CloseAudioDevice(SDL_AudioDeviceID dev);
void SDL_PauseAudioDevice(SDL_AudioDeviceID dev, int pause_on);
int SDL_QueueAudio(SDL_AudioDeviceID dev, const void *data, uint32_t len);
uint32_t SDL_GetQueuedAudioSize(SDL_AudioDeviceID dev);
void SDL_ClearQueuedAudio(SDL_AudioDeviceID dev);
void SDL_Delay(uint32_t ms);
void SDL_memset(void *dst, int c, unsigned long len);
#define SDL_zero(x) memset(&(x), 0, sizeof((x)))
#endif
