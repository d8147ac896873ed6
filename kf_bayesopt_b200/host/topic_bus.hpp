// topic_bus.hpp -- an in-process stand-in for the two ROS1 topics of the reference (ROS is not
// installed here): publishers/subscribers by topic name, a bounded per-subscriber callback queue
// (queue depth 10 like trial_node.cpp:17,150 and robot1d_bayesopt.cpp:36,39) and spinOnce().
// Message types are the field-for-field stand-ins of kfbo_trial.hpp.
#pragma once
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace kfbo {

class TopicBus {
public:
    template <class M>
    void Subscribe(const std::string &topic, size_t queue, std::function<void(const M &)> cb)
    {
        subs_[topic].push_back(Sub{queue, [cb](const void *m) { cb(*static_cast<const M *>(m)); }});
    }
    // Copies the message into every subscriber's queue (oldest dropped when full, like roscpp).
    template <class M>
    void Publish(const std::string &topic, const M &m)
    {
        auto it = subs_.find(topic);
        if (it == subs_.end()) return;
        for (size_t i = 0; i < it->second.size(); ++i) {
            Sub &s = it->second[i];
            if (pending_.size() >= 4096) pending_.pop_front();
            size_t queued = 0;
            for (auto &p : pending_) queued += (p.topic == topic && p.sub == i);
            if (queued >= s.queue) {
                for (auto p = pending_.begin(); p != pending_.end(); ++p)
                    if (p->topic == topic && p->sub == i) { pending_.erase(p); break; }
            }
            auto copy = std::make_shared<M>(m);
            auto call = s.call;   // by value: later Subscribe() calls may move the subscriber table
            pending_.push_back(Pending{topic, i, [copy, call]() { call(copy.get()); }});
        }
    }
    // Runs the callbacks queued so far (callbacks may publish; those run on the next spinOnce).
    size_t SpinOnce()
    {
        std::deque<Pending> now;
        now.swap(pending_);
        for (auto &p : now) p.run();
        return now.size();
    }

private:
    struct Sub { size_t queue; std::function<void(const void *)> call; };
    struct Pending { std::string topic; size_t sub; std::function<void()> run; };
    std::map<std::string, std::vector<Sub>> subs_;
    std::deque<Pending> pending_;
};

}  // namespace kfbo
