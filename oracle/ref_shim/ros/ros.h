// ros/ros.h -- stand-in so that the reference's read_para_vehicle.{h,cpp}, read_para_bayes.{h,cpp} and
// trial_node.cpp compile without ROS (test infrastructure, see Eigen/Dense in this directory).
//
// What it restates of roscpp's published behaviour, and nothing more:
//   * the parameter server: a process-wide name -> value map that oracle/ref_driver.cpp fills (from the
//     reference's own yaml files, parsed by the tests); NodeHandle::getParam returns false for a missing name or
//     a type it cannot convert (an int parameter reads as double, like XmlRpc ints do in roscpp);
//   * topics: advertise<T>() returns a Publisher whose publish() appends the message to the topic's outbox;
//     subscribe() registers a member-function callback; ros::spin() delivers every message queued in the
//     topic's inbox to the callbacks, in order, on the calling thread, and returns (roscpp's spin() returns on
//     shutdown; here "the inbox is empty" is the shutdown);
//   * BUS MODE (ShimMaster::set_bus(true); tests/native/boloop): two nodes of this process -- one thread each -- talk to each
//     other: publish() also queues the message for the process' own subscribers, spinOnce() delivers what is pending for the
//     callbacks THE CALLING THREAD registered (a node's callbacks run on its own thread, as in roscpp), spin() does that until
//     ros::shutdown().  The master is then shared by two threads: every access goes through its mutex.
// In- and outboxes hold type-erased shared_ptr<void>; the driver knows the message types of the two topics of
// the contract ("noise": std_msgs::Float64MultiArray, "cost": std_msgs::Float64).
#pragma once
#include <algorithm>   // roscpp's headers pull it in; trial_node.cpp relies on that for max_element
#include <atomic>
#include <chrono>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>      // roscpp's headers pull it in; src/test_trial.cpp relies on that for usleep

// roscpp is built on boost: callers that subscribe with a callable spell its type boost::function<void(const M::ConstPtr &)>
// (kf_bayesopt_b200/host/robot1d_kf.cpp under -DKFBO_WITH_ROS does).  Boost is not installed; the std equivalent stands in.
namespace boost {
template <typename Signature> using function = std::function<Signature>;
}

namespace ros {

struct ShimParam {
    enum Kind { kInt, kDouble, kString, kDoubleVector } kind = kInt;
    long long i = 0;
    double d = 0.0;
    std::string s;
    std::vector<double> v;
};

struct ShimCallback {
    std::thread::id owner;                                   // the thread (node) that subscribed
    std::function<void(const std::shared_ptr<void> &)> fn;
};

struct ShimMaster {
    std::map<std::string, ShimParam> params;
    std::map<std::string, std::deque<std::shared_ptr<void>>> inbox, outbox;
    std::map<std::string, std::vector<ShimCallback>> callbacks;
    std::mutex mu;
    bool bus = false;
    std::atomic<bool> down{false};
    static ShimMaster &get()
    {
        static ShimMaster m;
        return m;
    }
    void set_bus(bool on)
    {
        std::lock_guard<std::mutex> lk(mu);
        bus = on;
        down.store(false);
    }
    size_t subscribers(const std::string &topic)
    {
        std::lock_guard<std::mutex> lk(mu);
        const auto it = callbacks.find(topic);
        return it == callbacks.end() ? 0 : it->second.size();
    }
    // deliver what is queued for the callbacks of `who` (all callbacks: every = true); returns the number of deliveries
    size_t dispatch(std::thread::id who, bool every)
    {
        std::vector<std::pair<std::function<void(const std::shared_ptr<void> &)>, std::shared_ptr<void>>> work;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (auto &kv : callbacks) {
                std::vector<const ShimCallback *> mine;
                for (const ShimCallback &cb : kv.second)
                    if (every || cb.owner == who) mine.push_back(&cb);
                if (mine.empty()) continue;
                std::deque<std::shared_ptr<void>> &q = inbox[kv.first];
                while (!q.empty()) {
                    for (const ShimCallback *cb : mine) work.emplace_back(cb->fn, q.front());
                    q.pop_front();
                }
            }
        }
        for (auto &w : work) w.first(w.second);              // outside the lock: a callback publishes
        return work.size();
    }
    void drop_callbacks(std::thread::id who, bool every)
    {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &kv : callbacks) {
            std::vector<ShimCallback> keep;
            for (ShimCallback &cb : kv.second)
                if (!every && cb.owner != who) keep.push_back(cb);
            kv.second.swap(keep);
        }
    }
};

class Publisher {
public:
    Publisher() {}
    explicit Publisher(const std::string &topic) : topic_(topic) {}
    template <typename M> void publish(const M &msg) const
    {
        ShimMaster &m = ShimMaster::get();
        const std::shared_ptr<void> p = std::make_shared<M>(msg);
        std::lock_guard<std::mutex> lk(m.mu);
        m.outbox[topic_].push_back(p);
        if (m.bus) m.inbox[topic_].push_back(p);
    }
private:
    std::string topic_;
};

class Subscriber {};

class NodeHandle {
public:
    bool getParam(const std::string &name, int &out) const
    {
        const ShimParam *p = find(name);
        if (!p || p->kind != ShimParam::kInt) return false;
        out = (int)p->i;
        return true;
    }
    bool getParam(const std::string &name, double &out) const
    {
        const ShimParam *p = find(name);
        if (!p || (p->kind != ShimParam::kDouble && p->kind != ShimParam::kInt)) return false;
        out = p->kind == ShimParam::kDouble ? p->d : (double)p->i;
        return true;
    }
    bool getParam(const std::string &name, std::string &out) const
    {
        const ShimParam *p = find(name);
        if (!p || p->kind != ShimParam::kString) return false;
        out = p->s;
        return true;
    }
    bool getParam(const std::string &name, std::vector<double> &out) const
    {
        const ShimParam *p = find(name);
        if (!p || p->kind != ShimParam::kDoubleVector) return false;
        out = p->v;
        return true;
    }
    template <typename M> Publisher advertise(const std::string &topic, unsigned /*queue_size*/) { return Publisher(topic); }
    template <typename M, typename T>
    Subscriber subscribe(const std::string &topic, unsigned /*queue_size*/, void (T::*fp)(const std::shared_ptr<const M> &), T *obj)
    {
        ShimMaster &master = ShimMaster::get();
        std::lock_guard<std::mutex> lk(master.mu);
        master.callbacks[topic].push_back(ShimCallback{std::this_thread::get_id(), [fp, obj](const std::shared_ptr<void> &m) {
            const std::shared_ptr<const M> typed = std::static_pointer_cast<const M>(m);
            (obj->*fp)(typed);
        }});
        return Subscriber();
    }
    // subscribe<M>(topic, queue, boost::function<void(const M::ConstPtr &)>): roscpp's overload for callables
    template <typename M>
    Subscriber subscribe(const std::string &topic, unsigned /*queue_size*/, const std::function<void(const std::shared_ptr<const M> &)> &cb)
    {
        ShimMaster &master = ShimMaster::get();
        std::lock_guard<std::mutex> lk(master.mu);
        master.callbacks[topic].push_back(ShimCallback{std::this_thread::get_id(), [cb](const std::shared_ptr<void> &m) {
            const std::shared_ptr<const M> typed = std::static_pointer_cast<const M>(m);
            cb(typed);
        }});
        return Subscriber();
    }
private:
    static const ShimParam *find(const std::string &name)
    {
        const auto &ps = ShimMaster::get().params;
        const auto it = ps.find(name);
        return it == ps.end() ? nullptr : &it->second;
    }
};

namespace param {
// ros::param::get("~name", out): private names are looked up as given ("~vehicle_yaml" is the key the tests set)
inline bool get(const std::string &name, std::string &out) { return NodeHandle().getParam(name, out); }
}  // namespace param

inline void init(int &, char **, const std::string &) {}
inline bool ok() { return !ShimMaster::get().down.load(); }
inline void shutdown() { ShimMaster::get().down.store(true); }
inline void spinOnce() { ShimMaster::get().dispatch(std::this_thread::get_id(), false); }
inline void spin()
{
    ShimMaster &m = ShimMaster::get();
    bool bus;
    {
        std::lock_guard<std::mutex> lk(m.mu);
        bus = m.bus;
    }
    if (!bus) {
        while (m.dispatch(std::this_thread::get_id(), true) > 0) {}
        m.drop_callbacks(std::this_thread::get_id(), true);   // the node object the callbacks point into dies when its main() returns
        return;
    }
    while (ok())
        if (m.dispatch(std::this_thread::get_id(), false) == 0) std::this_thread::sleep_for(std::chrono::microseconds(20));
    m.drop_callbacks(std::this_thread::get_id(), false);
}
}  // namespace ros
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_FATAL(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
