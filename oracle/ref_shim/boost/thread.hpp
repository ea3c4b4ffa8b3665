/* TEST INFRASTRUCTURE shim (oracle/_ref build): the subset of Boost.Thread the reference engine uses
 * (FDTD/engine_multithread.cpp:58-200,313-401, FDTD/operator_multithread.cpp:118-141), on top of <thread>.
 * boost::barrier::wait(): blocks until `count` threads arrived, then resets (cyclic); like Boost's it is a
 * mutex + condition variable and an interruption point (thread_group::interrupt_all() is how the reference
 * stops its workers, engine_multithread.cpp:127,146). */
#pragma once
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <chrono>
#include <vector>
#include <memory>
namespace boost {
struct thread_interrupted {};
namespace detail_shim {
struct state { std::atomic<bool> interrupt; state() : interrupt(false) {} };
inline std::shared_ptr<state>& current() { static thread_local std::shared_ptr<state> s; return s; }
inline bool interruption_requested() { return current() && current()->interrupt.load(std::memory_order_relaxed); }
}
class thread {
public:
	typedef std::thread::id id;
	thread() {}
	template <class F> explicit thread(F f) : m_state(new detail_shim::state)
	{
		std::shared_ptr<detail_shim::state> st = m_state;
		m_t.reset(new std::thread([st, f]() mutable {
			detail_shim::current() = st;
			try { f(); } catch (const thread_interrupted&) {}
		}));
	}
	~thread() { if (m_t && m_t->joinable()) m_t->detach(); }
	void join() { if (m_t && m_t->joinable()) m_t->join(); }
	bool joinable() const { return m_t && m_t->joinable(); }
	void interrupt() { if (m_state) m_state->interrupt.store(true); }
	id get_id() const { return m_t ? m_t->get_id() : id(); }
	static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
private:
	thread(const thread&);
	thread& operator=(const thread&);
	std::shared_ptr<detail_shim::state> m_state;
	std::unique_ptr<std::thread> m_t;
};
namespace this_thread {
inline std::thread::id get_id() { return std::this_thread::get_id(); }
inline void yield() { std::this_thread::yield(); }
inline void interruption_point() { if (detail_shim::interruption_requested()) throw thread_interrupted(); }
}
class thread_group {
public:
	thread_group() {}
	~thread_group() { for (thread* t : m_threads) delete t; }
	void add_thread(thread* t) { m_threads.push_back(t); }
	template <class F> thread* create_thread(F f) { thread* t = new thread(f); m_threads.push_back(t); return t; }
	void join_all() { for (thread* t : m_threads) t->join(); }
	void interrupt_all() { for (thread* t : m_threads) t->interrupt(); }
	size_t size() const { return m_threads.size(); }
private:
	thread_group(const thread_group&);
	thread_group& operator=(const thread_group&);
	std::vector<thread*> m_threads;
};
class barrier {
public:
	explicit barrier(unsigned count) : m_threshold(count), m_count(count), m_generation(0) {}
	bool wait()
	{
		std::unique_lock<std::mutex> lk(m_mtx);
		unsigned gen = m_generation;
		if (--m_count == 0) {
			++m_generation;
			m_count = m_threshold;
			m_cv.notify_all();
			return true;
		}
		for (;;) {
			// the timeout only bounds how late an interrupt() is noticed; arrivals are signalled
			if (m_cv.wait_for(lk, std::chrono::milliseconds(20), [&] { return gen != m_generation; })) return false;
			if (detail_shim::interruption_requested()) throw thread_interrupted();
		}
	}
private:
	std::mutex m_mtx;
	std::condition_variable m_cv;
	unsigned m_threshold, m_count, m_generation;
};
typedef std::mutex mutex;
}
