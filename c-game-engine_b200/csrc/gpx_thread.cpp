// gpx_thread.cpp — the fixed-tick step loop around the physics update, host code only.
//
// Restates engine/src/subsystem/threads/PhysicsThread.c:59-159 without SDL: one "GamePhysics" thread, a control mutex
// (function pointer, quit flag, input queue) and a tick mutex held for the whole fixed update, `delta` = the previous
// tick's wall time (work + idle) in units of the 60 Hz target, clamped to the 10 ticks/s floor (Physics.h:12-22).
// Headless runs can pin `delta` to 1 and skip the sleep, which is what makes trajectories reproducible.
#include "../../include/gpx.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

constexpr double TARGET_NS = 1000000000.0 / 60.0;  // PHYSICS_TARGET_NS_D
constexpr double MIN_TPS_NS = 1000000000.0 / 10.0; // PHYSICS_MIN_NS_D

struct Ticker
{
	std::thread thread;
	std::mutex thread_mutex;  // physicsThreadMutex
	std::mutex tick_mutex;    // physicsTickMutex
	gpx_fixed_update_fn fn = nullptr;
	gpx_input_event_fn on_event = nullptr;
	void *state = nullptr;
	bool post_quit = false;
	int pinned = 0;
	std::vector<std::vector<unsigned char>> events;
	std::atomic<uint64_t> frame{0};         // GlobalState.physicsFrame
	std::atomic<uint64_t> last_tick_ns{0};  // what TickGraphUpdate receives (PhysicsThread.c:109)
	bool running = false;
};
Ticker g_t;

uint64_t now_ns()
{
	return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void sleep_precise(double ns)
{
	if (ns <= 0.0) return;  // the reference passes the unsigned difference on; an overrun simply does not sleep here
	const auto until = std::chrono::steady_clock::now() + std::chrono::nanoseconds((int64_t)ns);
	// sleep most of it, spin the last stretch (SDL_DelayPrecise does the same)
	if (ns > 2.0e6) std::this_thread::sleep_for(std::chrono::nanoseconds((int64_t)(ns - 1.5e6)));
	while (std::chrono::steady_clock::now() < until) std::this_thread::yield();
}

void thread_main()
{
	double last_tick_time = TARGET_NS;
	for (;;)
	{
		const uint64_t time_start = now_ns();
		g_t.thread_mutex.lock();
		g_t.tick_mutex.lock();
		if (g_t.post_quit)
		{
			g_t.thread_mutex.unlock();
			g_t.tick_mutex.unlock();
			return;
		}
		for (auto &e : g_t.events)
			if (g_t.on_event) g_t.on_event(g_t.state, e.data(), (uint64_t)e.size());
		g_t.events.clear();
		const int pinned = g_t.pinned;
		if (g_t.fn == nullptr)
		{
			g_t.frame++;
			g_t.thread_mutex.unlock();
			g_t.tick_mutex.unlock();
			sleep_precise(TARGET_NS);
			continue;
		}
		// the function is copied so the control mutex is free while it runs
		const gpx_fixed_update_fn update = g_t.fn;
		void *state = g_t.state;
		g_t.thread_mutex.unlock();

		// delta = the share of one tick the last one took, idle time included; about 1
		const double delta = pinned ? 1.0 : last_tick_time / TARGET_NS;
		update(state, delta);
		g_t.frame++;
		g_t.tick_mutex.unlock();

		uint64_t elapsed = now_ns() - time_start;
		if (!pinned) sleep_precise(TARGET_NS - (double)elapsed);
		elapsed = now_ns() - time_start;
		g_t.last_tick_ns = elapsed;
		last_tick_time = (double)elapsed < MIN_TPS_NS ? (double)elapsed : MIN_TPS_NS;
	}
}

}  // namespace

extern "C" {

int gpx_thread_init(void *state)
{
	if (g_t.running) return GPX_ERR_INVALID_ARG;
	g_t.fn = nullptr;
	g_t.on_event = nullptr;
	g_t.state = state;
	g_t.post_quit = false;
	g_t.pinned = 0;
	g_t.frame = 0;
	g_t.last_tick_ns = 0;
	g_t.events.clear();
	g_t.running = true;
	g_t.thread = std::thread(thread_main);
	return GPX_OK;
}

void gpx_thread_set_function(gpx_fixed_update_fn function)
{
	std::lock_guard<std::mutex> lk(g_t.thread_mutex);  // held by the loop until the running iteration has picked its function
	g_t.frame = 0;
	g_t.fn = function;
}

void gpx_thread_set_pinned_delta(int pinned)
{
	std::lock_guard<std::mutex> lk(g_t.thread_mutex);
	g_t.pinned = pinned;
}

void gpx_thread_set_input_handler(gpx_input_event_fn handler)
{
	std::lock_guard<std::mutex> lk(g_t.thread_mutex);
	g_t.on_event = handler;
}

void gpx_thread_queue_input_event(const void *event, uint64_t size)
{
	if (!event || !size) return;
	std::vector<unsigned char> copy((const unsigned char *)event, (const unsigned char *)event + size);
	std::lock_guard<std::mutex> lk(g_t.thread_mutex);
	g_t.events.push_back(std::move(copy));
}

void gpx_thread_terminate(void)
{
	if (!g_t.running) return;
	{
		std::lock_guard<std::mutex> lk(g_t.thread_mutex);
		g_t.post_quit = true;
	}
	g_t.thread.join();
	g_t.events.clear();
	g_t.running = false;
}

void gpx_thread_lock_tick_mutex(void) { g_t.tick_mutex.lock(); }
void gpx_thread_unlock_tick_mutex(void) { g_t.tick_mutex.unlock(); }
uint64_t gpx_thread_frame(void) { return g_t.frame; }
uint64_t gpx_thread_last_tick_ns(void) { return g_t.last_tick_ns; }

}  // extern "C"
